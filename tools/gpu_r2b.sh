#!/bin/bash
# round 2, pass B (2 GPUs): CUDA IPC feasibility (same GPU, GPU0<->GPU1) + 2-rank parity incl. the phase API
O=gpurun_out/r2b; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/ipc_probe tools/ipc_probe.cu > $O/ipc_probe.log 2>&1
timeout 120 /tmp/ipc_probe 2 >> $O/ipc_probe.log 2>&1; echo "exit $?" >> $O/ipc_probe.log
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > $O/pytest_mgpu.log 2>&1; echo "exit $?" >> $O/pytest_mgpu.log
cat $O/ipc_probe.log; tail -15 $O/pytest_mgpu.log
