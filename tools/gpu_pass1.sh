#!/bin/bash
# GPU pass: parity tests, A/B bench of the force-kernel variants, launch list + full ncu capture.  Usage: tools/gpu_pass1.sh <tag>
TAG=${1:-pass1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/pytest_parity.log 2>&1; echo "pytest exit $?" >> $O/pytest_parity.log
tail -5 $O/pytest_parity.log
MESO_PAIR_ONCE=0 timeout 300 python bench.py --steps 400 --warmup 50 --no-cpu-baseline > $O/bench_sp_twosided.json 2> $O/bench_sp_twosided.err
MESO_PAIR_ONCE=1 timeout 300 python bench.py --steps 400 --warmup 50 --no-cpu-baseline > $O/bench_sp_once.json 2> $O/bench_sp_once.err
MESO_PAIR_ONCE=1 timeout 300 python bench.py --steps 300 --warmup 50 --precision dp --case 48 --no-cpu-baseline > $O/bench_dp_once.json 2> $O/bench_dp_once.err
python - <<PY
import json
for f in ("bench_sp_twosided","bench_sp_once","bench_dp_once"):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], "roof %.3f"%d["roofline"]["frac"], d["phases"])
    except Exception as e: print(f, "FAILED", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_sp.csv \
    python tools/profile_step.py --case 64 --precision sp --steps 10 > $O/launches_sp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_dpd' -c 2 \
    -o $O/prof_sp python tools/profile_step.py --case 64 --precision sp --steps 3 > $O/prof_sp.log 2>&1
ls -la $O
