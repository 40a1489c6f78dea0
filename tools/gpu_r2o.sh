#!/bin/bash
# round 2, pass O (1 GPU): tile-staged neighbor build (TMA bulk copies of cell rows into shared memory)
O=gpurun_out/r2o; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_fixes.py -m gpu -x -q > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log
tail -15 $O/pytest.log
timeout 300 python bench.py --case 64 --no-cpu-baseline --steps 300 --warmup 50 > $O/bench_case64.json 2> $O/bench_case64.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_build_tiles' -c 1 \
    -o $O/prof_build python tools/profile_step.py --case 64 --precision sp --steps 10 > $O/prof_build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_sp.csv \
    python tools/profile_step.py --case 64 --precision sp --steps 10 > $O/launches_sp.log 2>&1
python - <<PY
import json
try:
    d=json.load(open('$O/bench_case64.json'))
    print('case64', '%.3e'%d['value'], 'e2e', d['e2e'] and '%.3e'%d['e2e']['value'], {k:(round(v['ms_total']/max(v['calls'],1),4),v['calls']) for k,v in d['phases'].items()}, 'frac %.3f'%d['roofline']['frac'])
except Exception as e:
    print('failed', e); print(open('$O/bench_case64.err').read()[-1500:])
PY
python tools/launch_summary.py $O/launches_sp.csv 2>/dev/null | head -14
