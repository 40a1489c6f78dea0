#!/bin/bash
# round 2, pass H (8 GPUs): multi-rank parity on real GPUs (2/4/8 ranks, fluid, polymer, phase API) + the north_star bench
O=gpurun_out/r2h; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "test_multi_gpu_parity" > $O/pytest_mgpu.log 2>&1; echo "exit $?" >> $O/pytest_mgpu.log
tail -15 $O/pytest_mgpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 --steps 100 --warmup 10 > $O/bench_200_n8.json 2> $O/bench_200_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29712 bench.py --gpus 8 --case 64 --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_case64_n8.json 2> $O/bench_case64_n8.err
timeout 600 $TR --nproc-per-node 2 --master-port 29713 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline > $O/bench_200_n2.json 2> $O/bench_200_n2.err
timeout 600 $TR --nproc-per-node 4 --master-port 29714 bench.py --gpus 4 --steps 60 --warmup 5 --no-cpu-baseline > $O/bench_200_n4.json 2> $O/bench_200_n4.err
python - <<PY
import json
for f in ('bench_200_n8','bench_case64_n8','bench_200_n2','bench_200_n4'):
    try:
        d=json.load(open('$O/%s.json'%f))
        print(f, '%.3e'%d['value'], 'e2e', d['e2e'] and '%.3e'%d['e2e']['value'], d['ms_per_step'], {k:(round(v['ms_total']/max(v['calls'],1),4),v['calls']) for k,v in d['phases'].items()}, d.get('parity_check'))
    except Exception as e:
        print(f, 'failed', e); print(open('$O/%s.err'%f).read()[-1500:])
PY
