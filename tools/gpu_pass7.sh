#!/bin/bash
TAG=${1:-pass7}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_lammps_deck.py -m gpu -q > $O/pytest_deck.log 2>&1; echo "pytest exit $?" >> $O/pytest_deck.log
tail -5 $O/pytest_deck.log
for v in 1 3 4; do
  MESO_NB_PER_ATOM=1 MESO_NB_VARIANT=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_nb$v.json 2> $O/bench_nb$v.err
done
timeout 600 python bench.py --case 200 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_200.json 2> $O/bench_200.err
python - <<PY
import json
for f in ("bench_nb1","bench_nb3","bench_nb4","bench_200"):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], "roof %.3f"%d["roofline"]["frac"], {k:(round(v["ms_total"]/max(v["calls"],1)*1e3,1)) for k,v in d["phases"].items()})
    except Exception as e: print(f, "FAILED", e, open("$O/%s.err"%f).read()[-600:])
PY
