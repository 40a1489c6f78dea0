#!/bin/bash
TAG=${1:-pass10}
N=${2:-2}
O=gpurun_out/$TAG
mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 400 --warmup 50 > $O/bench_n$N.json 2> $O/bench_n$N.err
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 200 --warmup 20 --case 100 > $O/bench_n${N}_c100.json 2> $O/bench_n${N}_c100.err
python - <<PY
import json
for f in ("bench_n$N","bench_n${N}_c100"):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], "roof %.3f"%d["roofline"]["frac"], {k:(round(v["ms_total"]/max(v["calls"],1)*1e3,1)) for k,v in d["phases"].items()})
    except Exception as e: print(f, "FAILED", e, open("$O/%s.err"%f).read()[-1500:])
PY
