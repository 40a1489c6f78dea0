"""Launches tests/mgpu_worker.py under torchrun: one rank per GPU where the box has enough GPUs, otherwise all ranks on
GPU 0 (the halo moves through CUDA IPC either way; NCCL only bootstraps in the first mode, gloo in the second)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpu():
    try:
        from meso_b200 import lib
        return lib.load().meso_device_count()
    except Exception:
        return 0


CASES = [(2, (12,), ()), (4, (12,), ()), (8, (12,), ()), (2, (6, 7, 12), ()), (2, (10,), ("--polymer",)), (8, (12,), ("--polymer",)),
         (2, (12,), ("--phases",)), (4, (12,), ("--phases",)), (2, (8,), ("--channel",)), (4, (8,), ("--channel",)), (8, (8,), ("--channel",))]


def case_id(c):
    return "%d-L%s%s" % (c[0], "x".join(str(v) for v in c[1]), "".join(e.replace("--", "-") for e in c[2]))


def run_worker(n, L, extra, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), "--L"] + [str(v) for v in L] + list(extra)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("mgpu parity OK") == 2


@pytest.mark.gpu
@pytest.mark.parametrize("n,L,extra", CASES, ids=[case_id(c) for c in CASES])
def test_multi_gpu_parity(n, L, extra):
    """one rank per GPU (NCCL bootstrap)"""
    if ngpu() < n:
        pytest.skip("needs %d GPUs" % n)
    run_worker(n, L, extra, 29500 + n)


ONE_GPU_CASES = [c for c in CASES if c[0] <= 4] + [(8, (12,), ())]


@pytest.mark.gpu
@pytest.mark.parametrize("n,L,extra", ONE_GPU_CASES, ids=[case_id(c) for c in ONE_GPU_CASES])
def test_multi_rank_parity_on_one_gpu(n, L, extra):
    """the same decompositions with every rank on GPU 0: runs on a single-GPU box (the driver's), exercises the complete
    multi-rank path -- migration, ghost creation in the reference's order, per-step refresh, bond migration, phase API"""
    if ngpu() < 1:
        pytest.skip("needs a GPU")
    run_worker(n, L, tuple(extra) + ("--share-gpu",), 29600 + n)
