"""ctypes front-end of the CPU oracle (oracle/meso_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline leg of bench.py -- never by meso_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libmeso_oracle.so")
_REFMATH = os.path.join(_HERE, "_ref", "libref_math.so")
LMP_SERIAL = os.path.join(_HERE, "_ref", "lmp_serial")


def build(force=False):
    """Compile the C restatement (and nothing else)."""
    src = os.path.join(_HERE, "meso_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/libmeso_oracle.so"])
    return _LIB


_lib = None
u32, i32, f64, f32 = C.c_uint32, C.c_int, C.c_double, C.c_float
P = C.POINTER


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    sig = {
        "orc_tea": (None, [P(u32), P(u32), i32]),
        "orc_premix_tea": (u32, [u32, u32, i32]),
        "orc_bit_space3": (u32, [u32]),
        "orc_interleave3": (u32, [u32, u32, u32]),
        "orc_mantissa": (u32, [f32, f32, f32]),
        "orc_brev": (u32, [u32]),
        "orc_seed_now": (u32, [u32, u32]),
        "orc_signature": (u32, [u32, i32, f32, f32, f32]),
        "orc_gaussian_sp": (f32, [u32, u32]),
        "orc_gaussian_dp": (f64, [u32, u32]),
        "orc_rsqrt": (f64, [f64]), "orc_sqrtd": (f64, [f64]), "orc_rcp": (f64, [f64]),
        "orc_log2d_frac": (f64, [f64]), "orc_exp2d_frac": (f64, [f64]),
        "orc_powd": (f64, [f64, f64]), "orc_sinpi": (f64, [f64]), "orc_cospi": (f64, [f64]),
        "orc_log2u": (f64, [u32]),
        "orc_world_create": (C.c_void_p, [P(f64), P(f64), P(i32), P(i32), i32, P(f64), P(f64),
                                          f64, f64, i32, u32, f64, i32]),
        "orc_world_destroy": (None, [C.c_void_p]),
        "orc_last_error": (C.c_char_p, []),
        "orc_world_set_atoms": (i32, [C.c_void_p, i32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
        "orc_world_setup": (i32, [C.c_void_p, i32, i32]),
        "orc_world_run": (i32, [C.c_void_p, i32, i32, i32]),
        "orc_initial_integrate": (None, [C.c_void_p, i32]),
        "orc_final_integrate": (None, [C.c_void_p, i32]),
        "orc_rebuild": (i32, [C.c_void_p]),
        "orc_forward_comm": (None, [C.c_void_p]),
        "orc_force_clear": (None, [C.c_void_p]),
        "orc_pack": (None, [C.c_void_p, u32]),
        "orc_pair_compute": (None, [C.c_void_p, i32, i32]),
        "orc_temperature": (f64, [C.c_void_p, i32]),
        "orc_set_timestep": (None, [C.c_void_p, C.c_long]),
        "orc_get_timestep": (C.c_long, [C.c_void_p]),
        "orc_nranks": (i32, [C.c_void_p]),
        "orc_counts": (None, [C.c_void_p, i32, P(i32), P(i32), P(i32), P(i32), P(i32)]),
        "orc_bins": (None, [C.c_void_p, i32, P(i32), P(f64), P(f64)]),
        "orc_get_atoms": (None, [C.c_void_p, i32] + [C.c_void_p] * 7),
        "orc_get_packed": (None, [C.c_void_p, i32, C.c_void_p, C.c_void_p]),
        "orc_get_virial": (None, [C.c_void_p, i32, C.c_void_p, C.c_void_p]),
        "orc_get_reorder": (None, [C.c_void_p, i32, C.c_void_p, C.c_void_p]),
        "orc_get_cells": (None, [C.c_void_p, i32, C.c_void_p, C.c_void_p]),
        "orc_get_stencil": (i32, [C.c_void_p, i32, i32, C.c_void_p]),
        "orc_get_neighbors": (None, [C.c_void_p, i32, C.c_void_p, C.c_void_p]),
        "orc_get_neighbors_transposed": (None, [C.c_void_p, i32, C.c_void_p]),
        "orc_world_set_bonds": (i32, [C.c_void_p, i32, i32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
        "orc_set_bond_coeff": (None, [C.c_void_p, i32, C.c_void_p, C.c_void_p]),
        "orc_set_special_lj12": (None, [C.c_void_p, f64]),
        "orc_bond_compute": (None, [C.c_void_p, i32, i32]),
        "orc_bond_energy": (f64, [C.c_void_p]),
        "orc_fix_add": (i32, [C.c_void_p, i32, i32, i32, P(f64)]),
        "orc_fix_clear": (None, [C.c_void_p]),
        "orc_fix_rdf_read": (i32, [C.c_void_p, i32, P(f64), P(f64), P(f64), P(f64)]),
        "orc_fix_post_force": (None, [C.c_void_p, i32]),
        "orc_fix_bounce": (None, [C.c_void_p, i32]),
        "orc_set_integrate_group": (None, [C.c_void_p, i32]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def ref_math():
    """The reference's own math_meso.h functions compiled host-side, or None
    (oracle/_ref/ is built in the dev container only; it travels to the GPU box)."""
    if not os.path.exists(_REFMATH):
        return None
    L = C.CDLL(_REFMATH)
    L.ref_tea.argtypes = [P(u32), P(u32), i32]
    for n in ("ref_premix16", "ref_premix64"):
        getattr(L, n).restype, getattr(L, n).argtypes = u32, [u32, u32]
    for n in ("ref_interleave3", "ref_morton"):
        getattr(L, n).restype, getattr(L, n).argtypes = u32, [u32, u32, u32]
    L.ref_mantissa.restype, L.ref_mantissa.argtypes = u32, [f32, f32, f32]
    L.ref_clamp.restype, L.ref_clamp.argtypes = i32, [i32, i32, i32]
    L.ref_gaussian_dp.restype, L.ref_gaussian_dp.argtypes = f64, [u32, u32]
    L.ref_gaussian_sp.restype, L.ref_gaussian_sp.argtypes = f32, [u32, u32]
    for n in ("ref_rsqrt", "ref_sqrtd", "ref_rcp", "ref_log2d_frac", "ref_exp2d_frac", "ref_sinpi", "ref_cospi"):
        getattr(L, n).restype, getattr(L, n).argtypes = f64, [f64]
    L.ref_powd.restype, L.ref_powd.argtypes = f64, [f64, f64]
    L.ref_log2u.restype, L.ref_log2u.argtypes = f64, [u32]
    return L


def tea(v0, v1, rounds):
    a, b = u32(v0), u32(v1)
    lib().orc_tea(C.byref(a), C.byref(b), rounds)
    return a.value, b.value


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_coeff(ntypes=1, a0=15.0, gamma=4.5, sigma=3.0, expw=1.0, cut=1.0):
    """[ntypes*ntypes][7] = {cut,cutsq,cutinv,expw,a0,gamma,sigma} (UM/pair_dpd_meso.h:15-24)."""
    row = [cut, cut * cut, 1.0 / cut, expw, a0, gamma, sigma]
    return np.tile(np.array(row, dtype=np.float64), (ntypes * ntypes, 1))


class World:
    """R simulated ranks of the reference algorithm in one process."""

    def __init__(self, boxlo, boxhi, periodic=(1, 1, 1), procgrid=(1, 1, 1), ntypes=1, mass=None,
                 coeff=None, cut_max=1.0, skin=0.3, every=5, seed=419084618, dt=0.005, precision=0):
        L = lib()
        mass = np.ascontiguousarray(mass if mass is not None else [0.0] + [1.0] * ntypes, dtype=np.float64)
        coeff = np.ascontiguousarray(coeff if coeff is not None else default_coeff(ntypes), dtype=np.float64)
        a3 = lambda v, t: (t * 3)(*v)
        self.h = L.orc_world_create(a3(boxlo, f64), a3(boxhi, f64), a3(periodic, i32), a3(procgrid, i32),
                                    ntypes, mass.ctypes.data_as(P(f64)), coeff.ctypes.data_as(P(f64)),
                                    cut_max, skin, every, seed & 0xFFFFFFFF, dt, precision)
        if not self.h:
            raise RuntimeError(L.orc_last_error().decode())
        self.L = L

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_world_destroy(self.h)
            self.h = None

    def _chk(self, rc):
        if rc:
            raise RuntimeError(self.L.orc_last_error().decode())

    def set_atoms(self, x, v=None, tag=None, type=None, mask=None, image=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = x.shape[0]
        c = lambda a, t: None if a is None else np.ascontiguousarray(a, dtype=t)
        v, tag, type, mask, image = c(v, np.float64), c(tag, np.int32), c(type, np.int32), c(mask, np.int32), c(image, np.int32)
        self._chk(self.L.orc_world_set_atoms(self.h, n, _p(x), _p(v), _p(tag), _p(type), _p(mask), _p(image)))

    def set_bonds(self, num_bond, bond_type, bond_atom, tag=None, k=None, r0=None, special_lj12=1.0):
        """per-atom bond tables (LAMMPS layout, same atom order as set_atoms) + bond_coeff arrays [nbondtypes+1]"""
        num_bond = np.ascontiguousarray(num_bond, np.int32)
        bond_type = np.ascontiguousarray(bond_type, np.int32).reshape(len(num_bond), -1)
        bond_atom = np.ascontiguousarray(bond_atom, np.int32).reshape(len(num_bond), -1)
        tag = None if tag is None else np.ascontiguousarray(tag, np.int32)
        self._chk(self.L.orc_world_set_bonds(self.h, len(num_bond), bond_type.shape[1], _p(tag), _p(num_bond), _p(bond_type), _p(bond_atom)))
        k, r0 = np.ascontiguousarray(k, np.float64), np.ascontiguousarray(r0, np.float64)
        self.L.orc_set_bond_coeff(self.h, len(k) - 1, _p(k), _p(r0))
        self.L.orc_set_special_lj12(self.h, float(special_lj12))

    def bond_compute(self, eflag=0, vflag=0): self.L.orc_bond_compute(self.h, eflag, vflag)
    def bond_energy(self): return self.L.orc_bond_energy(self.h)

    # channel fixes, argument meaning of the reference's constructors (UM/fix_wall_meso.cu:24-46 etc.)
    def _fix(self, kind, groupbit, dims, p):
        h = self.L.orc_fix_add(self.h, kind, groupbit, dims, (f64 * 4)(*(list(p) + [0.0] * (4 - len(p)))))
        if h < 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return h

    def fix_wall(self, dims, d, f, groupbit=1): return self._fix(1, groupbit, sum(1 << "xyz".index(a) for a in dims), [d, f])
    def fix_solid_bound(self, dims, groupbit=1): return self._fix(2, groupbit, sum(1 << "xyz".index(a) for a in dims), [])
    def fix_addforce(self, fx, fy, fz, groupbit=1): return self._fix(3, groupbit, 0, [fx, fy, fz])

    def fix_pois(self, dim_ortho, dim_force, strength, bisect_frac=0.5, groupbit=1):
        return self._fix(4, groupbit, dim_ortho | (dim_force << 2), [strength, bisect_frac])

    def fix_rdf(self, nbin, every=1, groupbit=1, other=None, rc=1.0):
        self._rdf_nbin = getattr(self, "_rdf_nbin", {})
        h = self._fix(5, groupbit, nbin, [every, groupbit if other is None else other, rc])
        self._rdf_nbin[h] = nbin
        return h

    def rdf(self, handle):
        """(histogram, samples, ni, nj)"""
        hist = np.zeros(self._rdf_nbin[handle])
        s, ni, nj = f64(), f64(), f64()
        self._chk(self.L.orc_fix_rdf_read(self.h, handle, hist.ctypes.data_as(P(f64)), C.byref(s), C.byref(ni), C.byref(nj)))
        return hist, int(s.value), ni.value, nj.value

    def fix_post_force(self, only=-1): self.L.orc_fix_post_force(self.h, only)
    def fix_bounce(self, only=-1): self.L.orc_fix_bounce(self.h, only)
    def integrate_group(self, groupbit): self.L.orc_set_integrate_group(self.h, groupbit)

    def setup(self, eflag=0, vflag=0):
        self._chk(self.L.orc_world_setup(self.h, eflag, vflag))

    def run(self, n, eflag=0, vflag=0):
        self._chk(self.L.orc_world_run(self.h, n, eflag, vflag))

    def rebuild(self):
        self._chk(self.L.orc_rebuild(self.h))

    def initial_integrate(self, groupbit=1): self.L.orc_initial_integrate(self.h, groupbit)
    def final_integrate(self, groupbit=1): self.L.orc_final_integrate(self.h, groupbit)
    def forward_comm(self): self.L.orc_forward_comm(self.h)
    def force_clear(self): self.L.orc_force_clear(self.h)
    def pack(self, seed_now): self.L.orc_pack(self.h, seed_now)
    def pair_compute(self, eflag=0, vflag=0): self.L.orc_pair_compute(self.h, eflag, vflag)
    def temperature(self, groupbit=1): return self.L.orc_temperature(self.h, groupbit)

    @property
    def ntimestep(self): return self.L.orc_get_timestep(self.h)

    @ntimestep.setter
    def ntimestep(self, t): self.L.orc_set_timestep(self.h, t)

    @property
    def nranks(self): return self.L.orc_nranks(self.h)

    def counts(self, r=0):
        v = [i32() for _ in range(5)]
        self.L.orc_counts(self.h, r, *[C.byref(a) for a in v])
        return dict(zip(("nlocal", "nghost", "n_bulk", "n_border", "n_col"), (a.value for a in v)))

    def bins(self, r=0):
        m, bs, bi = (i32 * 3)(), (f64 * 3)(), (f64 * 3)()
        self.L.orc_bins(self.h, r, m, bs, bi)
        return list(m), list(bs), list(bi)

    def atoms(self, r=0):
        c = self.counts(r)
        nl, n = c["nlocal"], c["nlocal"] + c["nghost"]
        out = dict(x=np.empty((n, 3)), v=np.empty((n, 3)), f=np.empty((nl, 3)), tag=np.empty(n, np.int32),
                   type=np.empty(n, np.int32), mask=np.empty(n, np.int32), image=np.empty(nl, np.int32))
        self.L.orc_get_atoms(self.h, r, *[_p(out[k]) for k in ("x", "v", "f", "tag", "type", "mask", "image")])
        out.update(c)
        return out

    def packed(self, r=0):
        c = self.counts(r)
        n = c["nlocal"] + c["nghost"]
        a, b = np.empty((n, 4), np.float32), np.empty((n, 4), np.float32)
        self.L.orc_get_packed(self.h, r, _p(a), _p(b))
        return a, b

    def virial(self, r=0):
        nl = self.counts(r)["nlocal"]
        a, b = np.empty((nl, 6)), np.empty(nl)
        self.L.orc_get_virial(self.h, r, _p(a), _p(b))
        return a, b

    def reorder(self, r=0):
        nl = self.counts(r)["nlocal"]
        k, p = np.empty(nl, np.uint64), np.empty(nl, np.int32)
        self.L.orc_get_reorder(self.h, r, _p(k), _p(p))
        return k, p

    def cells(self, r=0):
        c = self.counts(r)
        m, _, _ = self.bins(r)
        nc = m[0] * m[1] * m[2]
        s, a = np.empty(nc + 1, np.int32), np.empty(c["nlocal"] + c["nghost"], np.int32)
        self.L.orc_get_cells(self.h, r, _p(s), _p(a))
        return s, a

    def stencil(self, cell, r=0):
        o = np.empty(27, np.int32)
        n = self.L.orc_get_stencil(self.h, r, cell, _p(o))
        return o[:n].copy()

    def neighbors(self, r=0):
        c = self.counts(r)
        cnt, rows = np.empty(c["nlocal"], np.int32), np.empty((c["nlocal"], c["n_col"]), np.int32)
        self.L.orc_get_neighbors(self.h, r, _p(cnt), _p(rows))
        return cnt, rows

    def neighbors_transposed(self, r=0):
        c = self.counts(r)
        nrow = (c["nlocal"] + 31) // 32 * 32
        t = np.full(nrow * c["n_col"], -1, np.int32)
        self.L.orc_get_neighbors_transposed(self.h, r, _p(t))
        return t
