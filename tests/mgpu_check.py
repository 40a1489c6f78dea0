"""Multi-rank parity check of one brick of a decomposed box against the oracle's simulated-rank world.

Used by tests/mgpu_worker.py (one process per GPU under torchrun), by tests/test_multi_gpu.py and by the preflight of
bench.py (`parity_check` in its JSON line).  The oracle is the checker here, never the thing measured.

What is compared, per rank, with the oracle world of the same processor grid:
  * at setup (no migration yet): counts, local + ghost tags IN ORDER, fp64 ghost x,v bit for bit, ordered neighbor lists,
    packed coordinates / velocities / TEA signatures as uint32, forces (1e-5 fp32 / 1e-12 fp64), temperature;
  * fp64 style: a run that crosses rebuilds with migration -- ownership by tag, x and v to 1e-10, neighbor SETS by tag
    (the order of migrated atoms with identical sort keys is the one documented deviation), temperature;
  * phases=True: the same run spelled through the phase entry points (initial_integrate / rebuild | forward_comm /
    pair_compute / final_integrate), which exercises meso_forward_comm on a decomposition (ghost velocities and signatures
    must be this step's, or F_ij != -F_ji across brick faces).
"""
import numpy as np

import oracle
from meso_b200 import lib as _lib
from meso_b200 import workload
from meso_b200.engine import Meso


def brick_of(rank, grid):
    return (rank // (grid[1] * grid[2]), (rank // grid[2]) % grid[1], rank % grid[2])


AMPHI_COEFF = {(1, 1): 25.0, (2, 2): 25.0, (3, 3): 25.0, (1, 2): 27.0, (1, 3): 40.0, (2, 3): 40.0}


def make_inputs(dims, polymer=False, channel=False):
    """channel: BASELINE configs[4] -- amphiphilic bead-spring chains in solvent between two walls across z (solid_bound/meso),
    driven by pois/meso; three atom types, 1-2 exclusions; periodic in x and y only, so a grid that splits z puts the walls on a
    decomposed non-periodic dimension and chains across brick faces"""
    if channel:
        assert len(set(dims)) == 1
        x, typ, tag, nbond, btype, batom = workload.amphiphilic_channel(dims[0])
        coeff = np.zeros((3, 3, 7))
        for (a, b), a0 in AMPHI_COEFF.items():
            coeff[a - 1, b - 1] = coeff[b - 1, a - 1] = [1.0, 1.0, 1.0, 1.0, a0, 4.5, 3.0]
        v = workload.maxwell_velocities(len(x), seed=99) * 2.0
        return dict(x=x, v=v, tag=tag, typ=typ, ntypes=3, coeff=coeff.reshape(-1, 7), bonds=(nbond, btype, batom), polymer=True, channel=True)
    if polymer:
        assert len(set(dims)) == 1
        x, typ, tag, nbond, btype, batom = workload.polymer_melt(dims[0], chain_len=8, seed=5)
        ntypes = 2
        coeff = np.array([[1, 1, 1, 1, 25, 4.5, 3.0], [1, 1, 1, 1, 40, 4.5, 3.0], [1, 1, 1, 1, 40, 4.5, 3.0], [1, 1, 1, 1, 25, 4.5, 3.0]], float)
        bonds = (nbond, btype, batom)
    else:
        x = workload.dpd_fluid(dims if len(set(dims)) > 1 else dims[0])
        tag = np.arange(1, len(x) + 1, dtype=np.int32)
        typ, ntypes, coeff, bonds = np.ones(len(x), np.int32), 1, None, None
    v = workload.maxwell_velocities(len(x)) * (2.0 if polymer else 1.0)      # hotter chains: more migration within the run
    return dict(x=x, v=v, tag=tag, typ=typ, ntypes=ntypes, coeff=coeff, bonds=bonds, polymer=polymer, channel=False)


def check(precision, rank, grid, local_device, dims, inp, nccl_id, steps=12, phases=False, periodic=(1, 1, 1), dist=None):
    """returns the max relative force error at setup; raises AssertionError on any mismatch.  Collective: every rank of the
    grid calls it with the same arguments.  nccl_id shared by the ranks (one GPU per rank), or None: the ranks then exchange
    the library's halo blobs through `dist` (torch.distributed, any backend) and sum reductions themselves -- the mode in
    which several ranks share ONE GPU (CUDA IPC works between processes on the same device)."""
    host_boot = nccl_id is None

    def allsum(val):
        if not host_boot:
            return val
        vals = [None] * dist.get_world_size()
        dist.all_gather_object(vals, float(val))
        return sum(vals)
    x, v, tag, typ, ntypes, coeff, polymer = inp["x"], inp["v"], inp["tag"], inp["typ"], inp["ntypes"], inp["coeff"], inp["polymer"]
    channel = inp.get("channel", False)
    if channel:
        periodic = (1, 1, 0)
    w = oracle.World((0, 0, 0), dims, periodic=periodic, procgrid=grid, precision=1 if precision == "dp" else 0, ntypes=ntypes, coeff=coeff,
                     **({"mass": [0.0] + [1.0] * ntypes} if channel else {}))
    w.set_atoms(x, v, tag=tag, type=typ)
    if polymer:
        nbond, btype, batom = inp["bonds"]
        w.set_bonds(nbond, btype, batom, tag=tag, k=[0.0, 50.0], r0=[0.0, 0.5], special_lj12=0.0)
    if channel:
        w.fix_solid_bound("z"); w.fix_pois(2, 0, 0.2)
    w.setup(eflag=1, vflag=1)
    ao = w.atoms(rank)
    nl = ao["nlocal"]
    loc = brick_of(rank, grid)
    lo = np.array([dims[d] * (loc[d] * (1.0 / grid[d])) for d in range(3)])
    hi = np.array([dims[d] * ((loc[d] + 1) * (1.0 / grid[d])) if loc[d] < grid[d] - 1 else dims[d] for d in range(3)])
    mine = np.all((x >= lo) & (x < hi), axis=1)
    m = Meso(local_device)
    m.box((0.0, 0.0, 0.0), dims, periodic)
    m.decomposition(rank, grid, nccl_id)
    m.masses([0.0] + [1.0] * ntypes)
    m.neighbor(0.3, "bin")
    m.neigh_modify(delay=0, every=5, check=False)
    m.pair_style("dpd/fast/meso" if precision == "sp" else "dpd/meso", 1.0, 419084618)
    if channel:
        for (a, b), a0 in AMPHI_COEFF.items():
            m.pair_coeff(a, b, a0, 4.5, 3.0, 1.0, 1.0)
    elif polymer:
        m.pair_coeff(1, 1, 25, 4.5, 3.0, 1.0, 1.0); m.pair_coeff(1, 2, 40, 4.5, 3.0, 1.0, 1.0); m.pair_coeff(2, 2, 25, 4.5, 3.0, 1.0, 1.0)
    else:
        m.pair_coeff(1, 1, 15, 4.5, 3.0, 1.0, 1.0)
    m.timestep(0.005)
    m.upload(x[mine], v[mine], tag=tag[mine], type=typ[mine])
    if polymer:
        m.bond_style("harmonic/meso", 1)
        m.bond_coeff(1, 50.0, 0.5)
        m.special_bonds(0.0)
        m.bonds(nbond[mine], btype[mine], batom[mine], tag_max=len(x))
    if channel:
        m.fix("solid_bound/meso", "z", "rho5rc1s1"); m.fix("pois/meso", "z", "x", 0.2)
    if host_boot:
        blobs = [None] * dist.get_world_size()
        dist.all_gather_object(blobs, m.comm_export())
        m.comm_import(blobs)
    m.setup(eflag=1, vflag=1)
    cg, co = m.counts(), w.counts(rank)
    for k in ("nlocal", "nghost", "n_bulk", "n_border"):
        assert cg[k] == co[k], (rank, k, cg, co)
    ag = m.download()
    assert np.array_equal(ag["tag"], ao["tag"][:nl]), "local order differs"
    assert np.array_equal(ag["x"], ao["x"][:nl])
    gg = m.ghosts()
    assert np.array_equal(gg["tag"], ao["tag"][nl:]), "ghost order differs"
    assert np.array_equal(gg["x"], ao["x"][nl:]) and np.array_equal(gg["v"], ao["v"][nl:])
    cntg, rowsg = m.neighbors()
    cnto, rowso = w.neighbors(rank)
    msk = np.arange(rowso.shape[1])[None, :] < cnto[:, None]
    assert np.array_equal(cntg, cnto) and np.array_equal(rowsg[msk], rowso[msk]), "neighbor lists differ"
    c4g, v4g = m.packed()
    c4o, v4o = w.packed(rank)
    assert np.array_equal(c4g.view(np.uint32), c4o.view(np.uint32)) and np.array_equal(v4g.view(np.uint32), v4o.view(np.uint32))
    mag = np.linalg.norm(ao["f"], axis=1)
    err = (np.linalg.norm(ag["f"] - ao["f"], axis=1) / np.maximum(mag, mag.mean())).max()
    assert err <= ((2e-5 if polymer else 1e-5) if precision == "sp" else (1e-11 if polymer else 1e-12)), err
    if polymer:
        eb = allsum(m.bond_energy())                 # summed over the ranks by the library (collective call) or by the host
        assert abs(eb - w.bond_energy()) < 1e-10 * abs(w.bond_energy()), (eb, w.bond_energy())
    if not host_boot:
        assert m.L.meso_natoms_global(m.h) == len(x)

    def temperature():
        if not host_boot:
            return m.temperature()
        import ctypes as C
        s_, n_ = C.c_double(), C.c_double()
        m._chk(m.L.meso_compute_ke(m.h, 1, C.byref(s_), C.byref(n_)))
        return allsum(s_.value) / (3.0 * allsum(n_.value) - 3.0)

    t_g, t_o = temperature(), w.temperature()
    assert abs(t_g - t_o) < 1e-12, (t_g, t_o)

    def step_by_phases(mm):
        B, O, LOC = _lib.MESO_BULK, _lib.MESO_BORDER, _lib.MESO_LOCAL
        mm.ntimestep = mm.ntimestep + 1
        mm.initial_integrate()
        if mm.neighbor_decide():
            mm.rebuild()
            mm.force_clear(LOC)
            mm.pair_compute(B)
        else:
            mm.force_clear(B)
            mm.pair_compute(B)
            mm.forward_comm()
            mm.force_clear(O)
        mm.pair_compute(O)
        if polymer:
            mm.bond_compute()
        mm.final_integrate()

    if precision == "dp":
        if phases:
            for _ in range(steps):
                step_by_phases(m)
        else:
            m.run(steps)
        w.run(steps)
        ag, ao = m.download(), w.atoms(rank)
        nl = ao["nlocal"]
        assert m.counts()["nlocal"] == nl, (m.counts(), nl)
        og, oo = np.argsort(ag["tag"]), np.argsort(ao["tag"][:nl])
        assert np.array_equal(ag["tag"][og], ao["tag"][:nl][oo]), "ownership differs after migration"
        assert np.abs(ag["x"][og] - ao["x"][:nl][oo]).max() < 1e-10
        assert np.abs(ag["v"][og] - ao["v"][:nl][oo]).max() < 1e-10
        gg = m.ghosts()
        tags_g = np.concatenate([ag["tag"], gg["tag"]])
        cntg, rowsg = m.neighbors()
        cnto, rowso = w.neighbors(rank)
        sets_g = {int(ag["tag"][i]): frozenset(tags_g[rowsg[i, :cntg[i]]].tolist()) for i in range(nl)}
        sets_o = {int(ao["tag"][i]): frozenset(ao["tag"][rowso[i, :cnto[i]]].tolist()) for i in range(nl)}
        assert sets_g == sets_o, "neighbor sets differ after migration"
        assert abs(temperature() - w.temperature()) < 1e-10
    else:
        if phases:
            for _ in range(steps):
                step_by_phases(m)
        else:
            m.run(steps)
        t = temperature()
        assert 0.5 < t < (6.0 if polymer else 3.0), t      # the polymer case starts at T = 4 (velocities doubled)
    m.close()
    return float(err)
