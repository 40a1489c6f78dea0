#!/bin/bash
# One single-GPU box pass.  Usage: tools/gpu_round.sh <tag> [quick]
#   quick: parity tests of the rebuild path + the case-64 bench line
#   full : every GPU test, smoke(), the three bench lines (200^3 default, case 64 sp, case 48 dp), the ncu launch list and
#          --set full captures of the force kernels (sp, dp) and of the tile build
TAG=${1:-run}; MODE=${2:-full}
O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], '%.3e' % d['value'], 'e2e', d['e2e'] and '%.3e' % d['e2e']['value'],
          {k: round(v['ms_total'] / max(v['calls'], 1), 4) for k, v in d['phases'].items()}, 'frac %.3f' % d['roofline']['frac'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
if [ "$MODE" = quick ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fixes.py -m gpu -x -q > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log
  tail -3 $O/pytest.log
  timeout 300 python bench.py --case 64 --no-cpu-baseline --steps 300 --warmup 50 > $O/bench_case64.json 2> $O/bench_case64.err; show $O/bench_case64.json
  exit 0
fi
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_200.json 2> $O/bench_200.err; show $O/bench_200.json
timeout 600 python bench.py --case 64 --steps 1000 --warmup 100 > $O/bench_case64.json 2> $O/bench_case64.err; show $O/bench_case64.json
timeout 600 python bench.py --case 48 --precision dp --steps 500 --warmup 50 --no-cpu-baseline > $O/bench_dp48.json 2> $O/bench_dp48.err; show $O/bench_dp48.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_sp.csv \
    python tools/profile_step.py --case 64 --precision sp --steps 10 > $O/launches_sp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_dpd_once|k_build_tiles' -c 3 \
    -o $O/prof_sp python tools/profile_step.py --case 64 --precision sp --steps 5 > $O/prof_sp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_dpd_once' -c 1 \
    -o $O/prof_dp python tools/profile_step.py --case 48 --precision dp --steps 5 > $O/prof_dp.log 2>&1
python tools/launch_summary.py $O/launches_sp.csv 2>/dev/null | head -12
