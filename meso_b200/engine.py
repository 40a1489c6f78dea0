"""Host-side mirror of the USER-MESO plugin surface over the C ABI.

`Meso` exposes the deck vocabulary of example/simple/sp.run|dp.run
(neighbor, neigh_modify, pair_style, pair_coeff, timestep, run, compute temp/meso)
with the reference's argument meaning and error strings, and the phase calls of
ModifiedVerlet::run (UM/mvv_meso.cu:243-425) for parity tests.  It holds no
physics: every call goes to libmeso_b200.so.  The C++ LAMMPS package in
lammps/USER-MESO-B200 binds the same entry points.
"""
import ctypes as C

import numpy as np

from . import lib as _lib

PAIR_STYLES = {"dpd/fast/meso": _lib.MESO_SP, "dpd/meso": _lib.MESO_DP}


class MesoError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Meso:
    def __init__(self, device=0):
        """device: one GPU index, or a sequence of them -- several GPUs (or several bricks on one GPU) behind one handle"""
        self.L = _lib.load()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            rc = self.L.meso_create_gang(C.byref(h), len(device), (C.c_int * len(device))(*device))
        else:
            rc = self.L.meso_create(C.byref(h), device)
        if rc:
            raise MesoError("meso_create failed (%d): %s" % (rc, self.L.meso_last_error(None).decode()))
        self.h = h
        self.ntypes = 0
        self._coeff = None
        self._setflag = None
        self._cut_global = None
        self.dt = 0.005

    def close(self):
        if getattr(self, "h", None):
            self.L.meso_destroy(self.h)
            self.h = None

    __del__ = close

    def _chk(self, rc):
        if rc < 0:
            raise MesoError("%s (code %d)" % (self.L.meso_last_error(self.h).decode(), rc))
        return rc

    # ---- deck commands -------------------------------------------------
    def box(self, lo, hi, periodic=(1, 1, 1)):
        self._chk(self.L.meso_set_box(self.h, (C.c_double * 3)(*lo), (C.c_double * 3)(*hi), (C.c_int * 3)(*periodic)))

    def decomposition(self, rank, procgrid, nccl_id=None):
        buf = None
        if nccl_id is not None:
            buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
        self._chk(self.L.meso_set_decomposition(self.h, rank, (C.c_int * 3)(*procgrid), buf))

    def comm_export(self):
        """this rank's halo blob (host-driven bootstrap: ranks of one process or of one GPU); call after upload()"""
        self._push_coeff()                      # the ghost cutoff (largest pair cutoff + skin) sizes the arena
        buf = (C.c_char * self.L.meso_comm_blob_size())()
        self._chk(self.L.meso_comm_export(self.h, buf))
        return bytes(buf)

    def comm_import(self, blobs):
        """blobs: every rank's comm_export() in rank order"""
        raw = b"".join(blobs)
        buf = (C.c_char * len(raw)).from_buffer_copy(raw)
        self._chk(self.L.meso_comm_import(self.h, buf, len(blobs)))

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        rc = _lib.load().meso_comm_unique_id(buf)
        if rc:
            raise MesoError("meso_comm_unique_id failed (%d)" % rc)
        return bytes(buf)

    def masses(self, mass):
        """mass: 1-based list [unused, m1, m2, ...] like Atom::mass."""
        mass = np.ascontiguousarray(mass, dtype=np.float64)
        self.ntypes = len(mass) - 1
        self._chk(self.L.meso_set_types(self.h, self.ntypes, mass.ctypes.data_as(C.POINTER(C.c_double))))
        self._coeff = np.zeros((self.ntypes, self.ntypes, 7))
        self._setflag = np.zeros((self.ntypes, self.ntypes), dtype=bool)

    def neighbor(self, skin, style="bin"):
        if style != "bin":
            raise MesoError("Illegal neighbor command")
        self._skin = skin
        self._chk(self.L.meso_set_neighbor(self.h, skin, getattr(self, "_every", 5)))

    def neigh_modify(self, delay=0, every=1, check=False):
        if delay != 0 or check:
            raise MesoError("neigh_modify: only delay 0 / check no are supported on the device path")
        self._every = every
        self._chk(self.L.meso_set_neighbor(self.h, getattr(self, "_skin", 0.3), every))

    def pair_style(self, style, *args):
        """pair_style dpd/meso|dpd/fast/meso cut_global seed (UM/pair_dpd_meso.cu:272-277)."""
        if style not in PAIR_STYLES or len(args) != 2:
            raise MesoError("Illegal pair_style command")
        self._cut_global = float(args[0])
        self._chk(self.L.meso_pair_dpd_settings(self.h, PAIR_STYLES[style], float(args[0]), int(args[1])))

    def pair_coeff(self, i, j, *args):
        """pair_coeff I J a0 gamma sigma expw [cut] (UM/pair_dpd_meso.cu:290-327); I,J may be '*' ranges."""
        if len(args) < 4 or len(args) > 5 or self._coeff is None or self._cut_global is None:
            raise MesoError("Incorrect args for pair coefficients")
        a0, gamma, sigma, expw = (float(a) for a in args[:4])
        cut = float(args[4]) if len(args) == 5 else self._cut_global

        def bounds(s):
            s = str(s)
            if s == "*":
                return 1, self.ntypes
            if "*" in s:
                a, b = s.split("*")
                return (int(a) if a else 1), (int(b) if b else self.ntypes)
            return int(s), int(s)

        (ilo, ihi), (jlo, jhi) = bounds(i), bounds(j)
        count = 0
        for a in range(ilo, ihi + 1):
            for b in range(max(jlo, a), jhi + 1):
                row = [cut, cut * cut, 1.0 / cut, expw, a0, gamma, sigma]
                self._coeff[a - 1, b - 1] = row
                self._coeff[b - 1, a - 1] = row         # init_one mirrors i,j -> j,i
                self._setflag[a - 1, b - 1] = self._setflag[b - 1, a - 1] = True
                count += 1
        if count == 0:
            raise MesoError("Incorrect args for pair coefficients")

    def timestep(self, dt):
        self.dt = dt
        self._chk(self.L.meso_set_timestep_size(self.h, dt))

    def _push_coeff(self):
        if self._setflag is None or not self._setflag.all():
            raise MesoError("All pair coeffs are not set")
        c = np.ascontiguousarray(self._coeff.reshape(-1), dtype=np.float64)
        self._chk(self.L.meso_pair_dpd_coeff(self.h, c.ctypes.data_as(C.POINTER(C.c_double))))

    # ---- atom store ------------------------------------------------------
    def upload(self, x, v=None, tag=None, type=None, mask=None, image=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        c = lambda a, t: None if a is None else np.ascontiguousarray(a, dtype=t)
        v, tag, type, mask, image = c(v, np.float64), c(tag, np.int32), c(type, np.int32), c(mask, np.int32), c(image, np.int32)
        self._chk(self.L.meso_atoms_upload(self.h, x.shape[0], _ptr(x), _ptr(v), _ptr(tag), _ptr(type), _ptr(mask), _ptr(image)))

    # ---- bead-spring topology: bond_style harmonic/meso, bond_coeff, special_bonds, Bonds section ------------
    def bond_style(self, style, nbondtypes):
        if style != "harmonic/meso":
            raise MesoError("Invalid bond style")
        self._bond_k = np.zeros(nbondtypes + 1)
        self._bond_r0 = np.zeros(nbondtypes + 1)
        self._bond_set = np.zeros(nbondtypes + 1, dtype=bool)

    def bond_coeff(self, n, k, r0):
        """bond_coeff N K r0: E = K (r - r0)^2 (UM/bond_harmonic_meso.cu, src/MOLECULE/bond_harmonic.cpp)"""
        if getattr(self, "_bond_k", None) is None or not 1 <= n < len(self._bond_k):
            raise MesoError("Incorrect args for bond coefficients")
        self._bond_k[n], self._bond_r0[n], self._bond_set[n] = k, r0, True
        if self._bond_set[1:].all():
            self._chk(self.L.meso_bond_harmonic_coeff(self.h, len(self._bond_k) - 1, self._bond_k.ctypes.data_as(C.POINTER(C.c_double)),
                                                      self._bond_r0.ctypes.data_as(C.POINTER(C.c_double))))

    def special_bonds(self, lj12):
        self._chk(self.L.meso_set_special_bonds(self.h, float(lj12)))

    def bonds(self, num_bond, bond_type, bond_atom, tag_max=None):
        """per-atom table in LAMMPS' layout (rows [atom][slot]; bond_atom = partner TAG), same atom order as upload()"""
        num_bond = np.ascontiguousarray(num_bond, np.int32)
        bond_type = np.ascontiguousarray(bond_type, np.int32).reshape(len(num_bond), -1)
        bond_atom = np.ascontiguousarray(bond_atom, np.int32).reshape(len(num_bond), -1)
        if tag_max is None:
            tag_max = int(max(len(num_bond), bond_atom.max(initial=0)))
        self._chk(self.L.meso_bonds_upload(self.h, len(num_bond), bond_type.shape[1], _ptr(num_bond), _ptr(bond_type), _ptr(bond_atom),
                                           int(tag_max)))

    def bond_compute(self, eflag=0, vflag=0): self._chk(self.L.meso_bond_compute(self.h, eflag, vflag))

    def bond_energy(self):
        e = C.c_double()
        self._chk(self.L.meso_compute_bond_energy(self.h, C.byref(e)))
        return e.value

    # ---- fix wall/meso | solid_bound/meso | addforce/meso | pois/meso: same argument grammar as the reference ----------
    def fix(self, style, *args, groupbit=1):
        """fix ID group <style> args... (the ID and group words are implied: groupbit selects the atoms).  Returns the handle."""
        args = [str(a) for a in args]
        narg = len(args) + 3                                          # the reference counts ID, group and style
        if style == "wall/meso":                                     # UM/fix_wall_meso.cu:24-46
            if narg < 6:
                raise MesoError("Illegal fix MesoFixWall command")
            d = f = 0.0
            dims, i = 0, 0
            while i < len(args):
                a = args[i]
                if a in ("d", "f"):
                    i += 1
                    if i >= len(args):
                        raise MesoError("Incomplete fix wall command after '%s'" % a)
                    if a == "d":
                        d = float(args[i])
                    else:
                        f = float(args[i])
                elif a in ("x", "y", "z"):
                    dims |= 1 << "xyz".index(a)
                i += 1
            if not dims:
                raise MesoError("Incomplete fix wall command: insufficient arguments")
            return self._chk(self.L.meso_fix_wall(self.h, groupbit, dims, d, f))
        if style == "solid_bound/meso":                              # UM/fix_solid_bound_meso.cu:24-45
            if narg < 4:
                raise MesoError("Illegal fix MesoFixSolidBound command")
            dims = sum(1 << "xyz".index(a) for a in set(args) if a in ("x", "y", "z"))
            if not dims:
                raise MesoError("Incomplete fix wall command: dimension unspecified")
            if "rho5rc1s1" not in args:
                raise MesoError("Incomplete fix wall command: force kernel unspecified")
            return self._chk(self.L.meso_fix_solid_bound(self.h, groupbit, dims, 1))
        if style == "addforce/meso":                                 # UM/fix_addforce_meso.cu:24-32
            if narg < 6:
                raise MesoError("Illegal fix addforce/meso command")
            return self._chk(self.L.meso_fix_addforce(self.h, groupbit, *(float(a) for a in args[:3])))
        if style == "pois/meso":                                     # UM/fix_poiseuille_meso.cu:24-42
            if narg < 6:
                raise MesoError("Illegal fix CUDAPoiseuille command")
            dim = lambda a: int(a) if a[0].isdigit() else "xyz".index(a[0])
            frac = float(args[3]) if len(args) > 3 else 0.5
            return self._chk(self.L.meso_fix_pois(self.h, groupbit, dim(args[0]), dim(args[1]), float(args[2]), frac))
        if style == "rdf/fast/meso":                                 # UM/fix_rdf_fast_meso.cu:42-75
            output, nbin, every, other, i = "", 0, 1, groupbit, 0
            while i < len(args):
                a = args[i]
                if a in ("output", "nbin", "every", "other"):
                    i += 1
                    if i >= len(args):
                        raise MesoError("Incomplete compute vprof command after '%s'" % a)
                    if a == "output":
                        output = args[i]
                    elif a == "nbin":
                        nbin = int(args[i])
                    elif a == "every":
                        every = int(args[i])
                    else:
                        other = int(args[i])                         # the mirror takes the other group's bit directly
                i += 1
            if output == "" or nbin == 0:
                raise MesoError("Incomplete compute rdf command: insufficient arguments")
            h = self._chk(self.L.meso_fix_rdf(self.h, groupbit, other, nbin, every))
            self._rdf = getattr(self, "_rdf", {})
            self._rdf[h] = (output, nbin)
            return h
        raise MesoError("Invalid fix style")

    def rdf(self, handle, volume):
        """(r, g(r), histogram, samples) with the normalisation of MesoFixRDFFast::dump (UM/fix_rdf_fast_meso.cu:182-219, pi = 3.1415)"""
        output, nbin = self._rdf[handle]
        hist = np.zeros(nbin)
        ns, ni, nj = C.c_double(), C.c_double(), C.c_double()
        self._chk(self.L.meso_fix_rdf_read(self.h, handle, nbin, hist.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ns), C.byref(ni), C.byref(nj)))
        rc = self._cut_global
        bin_sz = rc / nbin
        i = np.arange(nbin)
        freq = hist / max(ni.value, 1.0) / max(ns.value, 1.0)
        shell = 4.0 / 3.0 * 3.1415 * ((bin_sz * (i + 1)) ** 3 - (bin_sz * i) ** 3)
        g = freq / shell / (nj.value / volume)
        return (i + 0.5) * bin_sz, g, hist, int(ns.value)

    def rdf_dump(self, handle, volume):
        """the file the reference writes when the fix is destroyed: `r <tab> g(r) <tab>` per bin, 15 significant digits"""
        r, g, _, _ = self.rdf(handle, volume)
        with open(self._rdf[handle][0], "w") as f:
            for a, b in zip(r, g):
                f.write("%.15g\t%.15g\t\n" % (a, b))

    def unfix_all(self): self._chk(self.L.meso_fix_clear(self.h))
    def fix_post_force(self, handle=-1): self._chk(self.L.meso_fix_post_force(self.h, handle))
    def fix_bounce(self, handle=-1): self._chk(self.L.meso_fix_bounce(self.h, handle))

    def counts(self):
        v = [C.c_int() for _ in range(4)]
        self._chk(self.L.meso_counts(self.h, *[C.byref(a) for a in v]))
        return dict(zip(("nlocal", "nghost", "n_bulk", "n_border"), (a.value for a in v)))

    def download(self, fields=("x", "v", "f", "tag", "type", "mask", "image")):
        n = self.counts()["nlocal"]
        out = {}
        spec = dict(x=((n, 3), np.float64), v=((n, 3), np.float64), f=((n, 3), np.float64), tag=((n,), np.int32),
                    type=((n,), np.int32), mask=((n,), np.int32), image=((n,), np.int32))
        for k in fields:
            out[k] = np.empty(*spec[k])
        args = [_ptr(out.get(k)) for k in ("x", "v", "f", "tag", "type", "mask", "image")]
        self._chk(self.L.meso_atoms_download(self.h, n, *args))
        return out

    # ---- run -------------------------------------------------------------
    def setup(self, eflag=0, vflag=0):
        self._push_coeff()
        self._chk(self.L.meso_setup(self.h, eflag, vflag))

    def run(self, nsteps, groupbit=1):
        self._chk(self.L.meso_run(self.h, nsteps, groupbit))

    def sync(self):
        self._chk(self.L.meso_sync(self.h))

    @property
    def ntimestep(self):
        return self.L.meso_get_ntimestep(self.h)

    @ntimestep.setter
    def ntimestep(self, t):
        self._chk(self.L.meso_set_ntimestep(self.h, t))

    # ---- phases of ModifiedVerlet::run -----------------------------------
    def initial_integrate(self, groupbit=1): self._chk(self.L.meso_initial_integrate(self.h, groupbit))
    def final_integrate(self, groupbit=1): self._chk(self.L.meso_final_integrate(self.h, groupbit))
    def neighbor_decide(self): return self._chk(self.L.meso_neighbor_decide(self.h))
    def rebuild(self): self._push_coeff(); self._chk(self.L.meso_rebuild(self.h))
    def forward_comm(self): self._chk(self.L.meso_forward_comm(self.h))
    def force_clear(self, range=_lib.MESO_LOCAL, vflag=0): self._chk(self.L.meso_force_clear(self.h, range, vflag))
    def pair_compute(self, range=_lib.MESO_LOCAL, eflag=0, vflag=0): self._chk(self.L.meso_pair_compute(self.h, range, eflag, vflag))

    # ---- compute temp/meso (UM/compute_temp_meso.cu:40-101; lj units) -----
    def temperature(self, groupbit=1, extra_dof=3):
        s, n = C.c_double(), C.c_double()
        self._chk(self.L.meso_compute_ke(self.h, groupbit, C.byref(s), C.byref(n)))
        dof = 3.0 * n.value - extra_dof
        return s.value / dof if dof > 0 else 0.0

    def virial(self):
        v, e = (C.c_double * 6)(), C.c_double()
        self._chk(self.L.meso_compute_virial(self.h, v, C.byref(e)))
        return np.array(v[:]), e.value

    # ---- exports for parity tests ------------------------------------------
    def bins(self):
        m, bs, bi, nc = (C.c_int * 3)(), (C.c_double * 3)(), (C.c_double * 3)(), C.c_int()
        self._chk(self.L.meso_export_bins(self.h, m, bs, bi, C.byref(nc)))
        return list(m), list(bs), list(bi), nc.value

    def reorder(self):
        n = self.counts()["nlocal"]
        k, p = np.empty(n, np.uint64), np.empty(n, np.int32)
        self._chk(self.L.meso_export_reorder(self.h, n, _ptr(k), _ptr(p)))
        return k, p

    def packed(self):
        c = self.counts()
        n = c["nlocal"] + c["nghost"]
        a, b = np.empty((n, 4), np.float32), np.empty((n, 4), np.float32)
        self._chk(self.L.meso_export_packed(self.h, n, _ptr(a), _ptr(b)))
        return a, b

    def ghosts(self):
        n = self.counts()["nghost"]
        x, v, tag, typ = np.empty((n, 3)), np.empty((n, 3)), np.empty(n, np.int32), np.empty(n, np.int32)
        self._chk(self.L.meso_export_ghosts(self.h, n, _ptr(x), _ptr(v), _ptr(tag), _ptr(typ)))
        return dict(x=x, v=v, tag=tag, type=typ)

    def cells(self):
        c = self.counts()
        m = self.bins()[0]
        nc = m[0] * m[1] * m[2]
        s, a = np.empty(nc + 1, np.int32), np.empty(c["nlocal"] + c["nghost"], np.int32)
        self._chk(self.L.meso_export_cells(self.h, nc + 1, _ptr(s), len(a), _ptr(a)))
        return s, a

    def stencil(self, cell):
        o = np.empty(27, np.int32)
        n = self._chk(self.L.meso_export_stencil(self.h, cell, _ptr(o)))
        return o[:n].copy()

    def pair_count(self):
        n = self.counts()["nlocal"]
        c = np.empty(n, np.int32)
        self._chk(self.L.meso_export_pair_count(self.h, n, _ptr(c)))
        return c

    def pair_table(self):
        n = self.counts()["nlocal"]
        n_col = self.bins()[3]
        t = np.empty(((n + 31) // 32 * 32) * n_col, np.int32)
        self._chk(self.L.meso_export_pair_table(self.h, t.size, _ptr(t)))
        return t, n_col

    def neighbors(self):
        """(pair_count, rows[nlocal][n_col]) de-transposed from the tile layout (UM/neigh_list_meso.cu:97-102)."""
        cnt = self.pair_count()
        t, n_col = self.pair_table()
        n = len(cnt)
        rows = np.full((n, n_col), -1, np.int32)
        i = np.arange(n)
        for k in range(int(cnt.max()) if n else 0):
            sel = cnt > k
            idx = ((i[sel] & ~31) + (k & 31)).astype(np.int64) * n_col + (k >> 5) * 32 + (i[sel] & 31)
            rows[sel, k] = t[idx]
        return cnt, rows

    def pair_rows(self):
        """production rows (pair_count, owned_count, rows[nlocal][n_col]): the layout the force kernels read"""
        cnt = self.pair_count()
        n = len(cnt)
        n_col = self.bins()[3]
        t = np.empty(((n + 31) // 32 * 32) * n_col, np.int32)
        own = np.empty(n, np.int32)
        self._chk(self.L.meso_export_pair_rows(self.h, t.size, _ptr(t), _ptr(own)))
        rows = np.full((n, n_col), -1, np.int32)
        i = np.arange(n)
        for k in range(int(cnt.max()) if n else 0):
            sel = cnt > k
            idx = ((i[sel] & ~31) + (k & 31)).astype(np.int64) * n_col + (k >> 5) * 32 + (i[sel] & 31)
            rows[sel, k] = t[idx]
        return cnt, own, rows

    def per_atom_virial(self):
        n = self.counts()["nlocal"]
        v, e = np.empty((n, 6)), np.empty(n)
        self._chk(self.L.meso_export_virial(self.h, n, _ptr(v), _ptr(e)))
        return v, e

    def eval_gaussian(self, si, sj):
        si, sj = np.ascontiguousarray(si, np.uint32), np.ascontiguousarray(sj, np.uint32)
        sp, dp = np.empty(len(si), np.float32), np.empty(len(si), np.float64)
        self._chk(self.L.meso_eval_gaussian(self.h, len(si), _ptr(si), _ptr(sj), _ptr(sp), _ptr(dp)))
        return sp, dp

    def eval_math(self, fn, a, b=None):
        names = ("rsqrt", "rcp", "log2d_frac", "exp2d_frac", "sinpi", "cospi", "sqrtd", "powd")
        a = np.ascontiguousarray(a, np.float64)
        b = None if b is None else np.ascontiguousarray(b, np.float64)
        out = np.empty_like(a)
        self._chk(self.L.meso_eval_math(self.h, names.index(fn), len(a), _ptr(a), _ptr(b), _ptr(out)))
        return out

    def eval_log2u(self, a):
        a = np.ascontiguousarray(a, np.uint32)
        out = np.empty(len(a), np.float64)
        self._chk(self.L.meso_eval_log2u(self.h, len(a), _ptr(a), _ptr(out)))
        return out

    # ---- timers --------------------------------------------------------------
    def timers(self, enable=None, reset=False):
        if enable is not None:
            self._chk(self.L.meso_timers_enable(self.h, int(enable)))
        ms, calls = (C.c_double * 5)(), (C.c_int64 * 5)()
        self._chk(self.L.meso_timers_read(self.h, ms, calls, int(reset)))
        return {n: (ms[i], calls[i]) for i, n in enumerate(_lib.TIMER_NAMES)}

    def launch_count(self, reset=False):
        n = C.c_int64()
        self._chk(self.L.meso_launch_count(self.h, C.byref(n), int(reset)))
        return n.value

    def stream(self):
        return self.L.meso_stream(self.h)

    def memory_usage(self):
        b = C.c_uint64()
        self._chk(self.L.meso_memory_usage(self.h, C.byref(b)))
        return b.value


def dpd_fluid_deck(L, precision="sp", device=0, x=None, v=None, seed=419084618):
    """The configuration of example/simple/sp.run|dp.run on a box of edge L (synthetic data if x is None)."""
    from . import workload
    if x is None:
        x = workload.dpd_fluid(L)
    if v is None:
        v = workload.maxwell_velocities(len(x))
    dims = (L, L, L) if np.isscalar(L) else L
    m = Meso(device)
    m.box((0.0, 0.0, 0.0), dims)
    m.masses([0.0, 1.0])
    m.neighbor(0.3, "bin")
    m.neigh_modify(delay=0, every=5, check=False)
    m.pair_style("dpd/fast/meso" if precision == "sp" else "dpd/meso", 1.0, seed)
    m.pair_coeff(1, 1, 15, 4.5, 3.0, 1.0, 1.0)
    m.timestep(0.005)
    m.upload(x, v)
    return m
