"""World-size-2 CPU test (gloo) of the host-side logic of the multi-GPU path: the processor grid bench.py picks, the brick
each rank claims, the rank order the library expects ((ix*py+iy)*pz+iz, include/meso_b200.h) and the weak-scaling tag layout.
The device side of the same path is covered on GPUs by tests/test_multi_gpu.py."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %r)
from bench import procgrid_for
from meso_b200 import workload
import oracle

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
grid = procgrid_for(world)
assert grid[0] * grid[1] * grid[2] == world
L = 6
dims = tuple(g * L for g in grid)
loc = (rank // (grid[1] * grid[2]), (rank // grid[2]) %% grid[1], rank %% grid[2])
assert (loc[0] * grid[1] + loc[1]) * grid[2] + loc[2] == rank
# weak-scaling workload of bench.py: every rank generates its own brick
x = workload.dpd_fluid(L, seed=workload.DEFAULT_SEED + rank) + np.array([loc[0] * L, loc[1] * L, loc[2] * L], dtype=np.float64)
tag = (np.arange(len(x), dtype=np.int64) + 1 + rank * len(x)).astype(np.int32)
# the oracle's simulated-rank world must assign exactly these atoms to this rank
allx = [None] * world
alltag = [None] * world
dist.all_gather_object(allx, x)
dist.all_gather_object(alltag, tag)
gx, gt = np.concatenate(allx), np.concatenate(alltag)
assert len(set(gt.tolist())) == len(gt)
w = oracle.World((0, 0, 0), dims, procgrid=grid)
w.set_atoms(gx, np.zeros_like(gx), tag=gt)
w.setup()
mine = w.atoms(rank)
assert mine["nlocal"] == len(x), (rank, mine["nlocal"], len(x))
assert set(mine["tag"][:mine["nlocal"]].tolist()) == set(tag.tolist())
counts = [None] * world
dist.all_gather_object(counts, w.counts(rank))
assert sum(c["nlocal"] for c in counts) == len(gx)
# halo volume sanity: every rank of a 2-rank split sees ghosts from its partner and from its own periodic images
assert counts[rank]["nghost"] > 0 and counts[rank]["n_border"] > 0
dist.barrier()
if rank == 0:
    print("gloo decomposition OK", grid, dims)
dist.destroy_process_group()
"""


def test_two_rank_decomposition_host_logic(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "gloo decomposition OK" in out.stdout
