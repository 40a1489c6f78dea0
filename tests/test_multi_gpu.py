"""Launches tests/mgpu_worker.py under torchrun on 2 (and 4/8 if present) GPUs of this node."""
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


@pytest.mark.gpu
@pytest.mark.parametrize("n,L,extra", [(2, (12,), ()), (4, (12,), ()), (8, (12,), ()), (2, (6, 7, 12), ()), (2, (10,), ("--polymer",)),
                                       (8, (12,), ("--polymer",)), (2, (12,), ("--phases",)), (4, (12,), ("--phases",))])
def test_multi_gpu_parity(n, L, extra):
    if ngpu() < n:
        pytest.skip("needs %d GPUs" % n)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + n), os.path.join(ROOT, "tests", "mgpu_worker.py"), "--L"] + [str(v) for v in L] + list(extra)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("mgpu parity OK") == 2
