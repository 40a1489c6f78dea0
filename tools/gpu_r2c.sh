#!/bin/bash
# round 2, pass C (1 GPU): the fine-lattice neighbor build + owned-prefix rows: parity, bench, ncu
O=gpurun_out/r2c; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 300 python bench.py --case 64 --no-cpu-baseline --steps 300 --warmup 50 > $O/bench_case64.json 2> $O/bench_case64.err
MESO_NB_SLOW=1 timeout 300 python bench.py --case 64 --no-cpu-baseline --no-e2e --no-parity --steps 100 --warmup 20 > $O/bench_case64_slow.json 2> $O/bench_case64_slow.err
timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > $O/bench_200.json 2> $O/bench_200.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_build_rows|k_dpd_once' -c 4 \
    -o $O/prof_sp python tools/profile_step.py --case 64 --precision sp --steps 5 > $O/prof_sp.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_sp.csv \
    python tools/profile_step.py --case 64 --precision sp --steps 10 > $O/launches_sp.log 2>&1
for f in bench_case64 bench_case64_slow bench_200; do python - <<PY
import json
try:
    d=json.load(open('$O/$f.json'))
    print('$f', '%.3e'%d['value'], 'e2e', d['e2e'] and '%.3e'%d['e2e']['value'], {k:(round(v['ms_total']/max(v['calls'],1),4),v['calls']) for k,v in d['phases'].items()}, 'frac %.3f'%d['roofline']['frac'], d.get('parity_check'))
except Exception as e:
    print('$f failed', e); print(open('$O/$f.err').read()[-1500:])
PY
done
