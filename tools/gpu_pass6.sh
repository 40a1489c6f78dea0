#!/bin/bash
TAG=${1:-pass6}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_cell.json 2> $O/bench_cell.err
for v in 0 1 2 3 4; do
  MESO_NB_PER_ATOM=1 MESO_NB_VARIANT=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_nb$v.json 2> $O/bench_nb$v.err
done
python - <<PY
import json
for f in ("bench_cell","bench_nb0","bench_nb1","bench_nb2","bench_nb3","bench_nb4"):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "roof %.3f"%d["roofline"]["frac"], {k:(round(v["ms_total"]/max(v["calls"],1)*1e3,1)) for k,v in d["phases"].items()})
    except Exception as e: print(f, "FAILED", e, open("$O/%s.err"%f).read()[-300:])
PY
