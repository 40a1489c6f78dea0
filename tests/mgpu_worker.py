"""Multi-GPU parity worker (one process per GPU, launched by torchrun):
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_worker.py [--L 12] [--grid px py pz]
Every rank owns one brick of the box; results are compared with the oracle's simulated-rank world of the same
processor grid: counts, local + ghost tags IN ORDER and ordered neighbor lists at setup (no migration yet), fp64
trajectories by tag after a run that crosses rebuilds with migration (neighbor SETS by tag there: the order of
migrated atoms with identical sort keys is the one documented deviation)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402
from meso_b200 import workload  # noqa: E402
from meso_b200.engine import Meso  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, nargs="+", default=[12])
ap.add_argument("--grid", type=int, nargs=3, default=None)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--backend", default="nccl")
ap.add_argument("--polymer", action="store_true", help="bead-spring chains in solvent (bond table rides the migration, ghost partners)")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group(a.backend, device_id=torch.device("cuda", local) if a.backend == "nccl" else None)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import procgrid_for  # noqa: E402

grid = tuple(a.grid) if a.grid else procgrid_for(world)
assert grid[0] * grid[1] * grid[2] == world
dims = tuple(a.L) * 3 if len(a.L) == 1 else tuple(a.L)
if a.polymer:
    assert len(set(dims)) == 1
    x, typ, tag, nbond, btype, batom = workload.polymer_melt(dims[0], chain_len=8, seed=5)
    ntypes, coeff = 2, np.array([[1, 1, 1, 1, 25, 4.5, 3.0], [1, 1, 1, 1, 40, 4.5, 3.0], [1, 1, 1, 1, 40, 4.5, 3.0], [1, 1, 1, 1, 25, 4.5, 3.0]], float)
else:
    x = workload.dpd_fluid(dims if len(set(dims)) > 1 else dims[0])
    tag = np.arange(1, len(x) + 1, dtype=np.int32)
    typ, ntypes, coeff = np.ones(len(x), np.int32), 1, None
v = workload.maxwell_velocities(len(x)) * (2.0 if a.polymer else 1.0)      # hotter chains: more migration within the run


def check(precision):
    w = oracle.World((0, 0, 0), dims, procgrid=grid, precision=1 if precision == "dp" else 0, ntypes=ntypes, coeff=coeff)
    w.set_atoms(x, v, tag=tag, type=typ)
    if a.polymer:
        w.set_bonds(nbond, btype, batom, tag=tag, k=[0.0, 50.0], r0=[0.0, 0.5], special_lj12=0.0)
    w.setup(eflag=1, vflag=1)
    ao = w.atoms(rank)
    nl = ao["nlocal"]
    # this rank's atoms, in the oracle's pre-sort (file) order: select by ownership from the global arrays
    loc = (rank // (grid[1] * grid[2]), (rank // grid[2]) % grid[1], rank % grid[2])
    lo = np.array([dims[d] * (loc[d] * (1.0 / grid[d])) for d in range(3)])
    hi = np.array([dims[d] * ((loc[d] + 1) * (1.0 / grid[d])) if loc[d] < grid[d] - 1 else dims[d] for d in range(3)])
    mine = np.all((x >= lo) & (x < hi), axis=1)
    m = Meso(local)
    m.box((0.0, 0.0, 0.0), dims)
    ids = [Meso.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    m.decomposition(rank, grid, ids[0])
    m.masses([0.0] + [1.0] * ntypes)
    m.neighbor(0.3, "bin")
    m.neigh_modify(delay=0, every=5, check=False)
    m.pair_style("dpd/fast/meso" if precision == "sp" else "dpd/meso", 1.0, 419084618)
    if a.polymer:
        m.pair_coeff(1, 1, 25, 4.5, 3.0, 1.0, 1.0); m.pair_coeff(1, 2, 40, 4.5, 3.0, 1.0, 1.0); m.pair_coeff(2, 2, 25, 4.5, 3.0, 1.0, 1.0)
    else:
        m.pair_coeff(1, 1, 15, 4.5, 3.0, 1.0, 1.0)
    m.timestep(0.005)
    m.upload(x[mine], v[mine], tag=tag[mine], type=typ[mine])
    if a.polymer:
        m.bond_style("harmonic/meso", 1)
        m.bond_coeff(1, 50.0, 0.5)
        m.special_bonds(0.0)
        m.bonds(nbond[mine], btype[mine], batom[mine], tag_max=len(x))
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
    assert err <= ((2e-5 if a.polymer else 1e-5) if precision == "sp" else (1e-11 if a.polymer else 1e-12)), err
    if a.polymer:
        eb = m.bond_energy()                         # summed over the ranks by the library (collective call)
        assert abs(eb - w.bond_energy()) < 1e-10 * abs(w.bond_energy()), (eb, w.bond_energy())
    assert m.L.meso_natoms_global(m.h) == len(x)
    t_g, t_o = m.temperature(), w.temperature()
    assert abs(t_g - t_o) < 1e-12, (t_g, t_o)
    if precision == "dp":
        m.run(a.steps)
        w.run(a.steps)
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
        assert abs(m.temperature() - w.temperature()) < 1e-10
    else:
        m.run(a.steps)
        t = m.temperature()
        assert 0.5 < t < (6.0 if a.polymer else 3.0), t      # the polymer case starts at T = 4 (velocities doubled)
    m.close()
    return err


for precision in ("dp", "sp"):
    e = check(precision)
    dist.barrier()
    if rank == 0:
        print("mgpu parity OK: %d ranks grid %s box %s %s%s (force err %.2e)" % (world, grid, dims, precision, " polymer" if a.polymer else "", e),
              flush=True)
dist.destroy_process_group()
