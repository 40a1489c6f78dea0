#!/bin/bash
TAG=${1:-pass4}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/pytest_parity.log 2>&1; echo "pytest exit $?" >> $O/pytest_parity.log
tail -5 $O/pytest_parity.log
for v in 0 2; do
  MESO_PAIR_TEX=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_tex$v.json 2> $O/bench_tex$v.err
done
MESO_NB_PER_ATOM=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_nbold.json 2> $O/bench_nbold.err
python - <<PY
import json
for f in ("bench_tex0","bench_tex2","bench_nbold"):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "roof %.3f"%d["roofline"]["frac"], {k:(round(v["ms_total"]/max(v["calls"],1)*1e3,1)) for k,v in d["phases"].items()})
    except Exception as e: print(f, "FAILED", e, open("$O/%s.err"%f).read()[-300:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_sp.csv \
    python tools/profile_step.py --case 64 --precision sp --steps 10 > $O/launches_sp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_build_neighbors_cell' -c 1 \
    -o $O/prof_nb python tools/profile_step.py --case 64 --precision sp --steps 6 > $O/prof_nb.log 2>&1
ls -la $O
