#!/bin/bash
TAG=${1:-pass13}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/pytest_parity.log 2>&1; echo "pytest exit $?" >> $O/pytest_parity.log
tail -15 $O/pytest_parity.log
timeout 300 python bench.py --steps 300 --warmup 50 --precision dp --case 48 --no-cpu-baseline > $O/bench_dp.json 2> $O/bench_dp.err
python - <<PY
import json
for f in ("bench_dp",):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "roof %.3f"%d["roofline"]["frac"], {k:(round(v["ms_total"]/max(v["calls"],1)*1e3,1)) for k,v in d["phases"].items()})
    except Exception as e: print(f, "FAILED", e, open("$O/%s.err"%f).read()[-600:])
PY
