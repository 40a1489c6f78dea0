#!/bin/bash
# 8-GPU pass: parity vs the simulated-rank oracle at 4 and 8 ranks, weak-scaling bench at N = 8 (case 64 and the 200^3 / 32M target)
TAG=${1:-pass11}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "8- or 4-" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 300 --warmup 30 > $O/bench_n8.json 2> $O/bench_n8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 200 --warmup 20 --case 100 > $O/bench_n8_c100.json 2> $O/bench_n8_c100.err
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 4 --steps 300 --warmup 30 > $O/bench_n4.json 2> $O/bench_n4.err
python - <<PY
import json
for f in ("bench_n8","bench_n8_c100","bench_n4"):
    try:
        txt=[l for l in open("$O/%s.json"%f) if l.startswith("{")][0]
        d=json.loads(txt); print(f, "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], "roof %.3f"%d["roofline"]["frac"], {k:(round(v["ms_total"]/max(v["calls"],1)*1e3,1)) for k,v in d["phases"].items()}, d["config"]["procgrid"])
    except Exception as e: print(f, "FAILED", e, open("$O/%s.err"%f).read()[-1500:])
PY
