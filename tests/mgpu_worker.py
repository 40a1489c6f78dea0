"""Multi-GPU parity worker (one process per GPU, launched by torchrun):
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_worker.py [--L 12] [--grid px py pz]
                                                                              [--polymer] [--phases]
Every rank owns one brick of the box; results are compared with the oracle's simulated-rank world of the same
processor grid by tests/mgpu_check.py (what is compared is listed there)."""
import argparse
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import mgpu_check  # noqa: E402
from bench import procgrid_for  # noqa: E402
from meso_b200.engine import Meso  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, nargs="+", default=[12])
ap.add_argument("--grid", type=int, nargs=3, default=None)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--backend", default="nccl")
ap.add_argument("--polymer", action="store_true", help="bead-spring chains in solvent (bond table rides the migration, ghost partners)")
ap.add_argument("--channel", action="store_true", help="BASELINE configs[4]: amphiphilic chains between walls across z, body force; periodic in x, y only")
ap.add_argument("--phases", action="store_true", help="drive the run through the phase entry points (meso_forward_comm on a decomposition)")
ap.add_argument("--share-gpu", action="store_true", help="all ranks on GPU 0: halo through CUDA IPC on one device, bootstrap and reductions through gloo")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
if a.share_gpu:
    local, a.backend = 0, "gloo"
torch.cuda.set_device(local)
dist.init_process_group(a.backend, device_id=torch.device("cuda", local) if a.backend == "nccl" else None)

grid = tuple(a.grid) if a.grid else procgrid_for(world)
assert grid[0] * grid[1] * grid[2] == world
dims = tuple(a.L) * 3 if len(a.L) == 1 else tuple(a.L)
inp = mgpu_check.make_inputs(dims, a.polymer, a.channel)

for precision in ("dp", "sp"):
    ids = [None]
    if not a.share_gpu:
        ids = [Meso.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
    e = mgpu_check.check(precision, rank, grid, local, dims, inp, ids[0], steps=a.steps, phases=a.phases, dist=dist)
    dist.barrier()
    if rank == 0:
        print("mgpu parity OK: %d ranks%s grid %s box %s %s%s%s (force err %.2e)" % (world, " on one GPU" if a.share_gpu else "", grid, dims, precision,
                                                                                   " channel" if a.channel else " polymer" if a.polymer else "", " phases" if a.phases else "", e), flush=True)
dist.destroy_process_group()
