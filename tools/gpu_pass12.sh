#!/bin/bash
TAG=${1:-pass12}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -q -k "2- or halo or pair_once" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 400 --warmup 50 > $O/bench_n2.json 2> $O/bench_n2.err
python - <<PY
import json
for f in ("bench_n2",):
    try:
        txt=[l for l in open("$O/%s.json"%f) if l.startswith("{")][0]
        d=json.loads(txt); print(f, "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], "roof %.3f"%d["roofline"]["frac"], {k:(round(v["ms_total"]/max(v["calls"],1)*1e3,1)) for k,v in d["phases"].items()}, d["config"]["procgrid"])
    except Exception as e: print(f, "FAILED", e, open("$O/%s.err"%f).read()[-1500:])
PY
head -c 300 $O/bench_n2.json
