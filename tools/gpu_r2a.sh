#!/bin/bash
# round 2, pass A (1 GPU): paths that round 1 left unrun + dp capture + 200^3 on one GPU
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
MESO_NB_SKIP=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fixes.py -m gpu -x -q > $O/nbskip_pytest.log 2>&1; echo "exit $?" >> $O/nbskip_pytest.log
for s in 0 1; do MESO_NB_SKIP=$s timeout 300 python bench.py --no-cpu-baseline --steps 300 --warmup 50 > $O/bench_nbskip$s.json 2> $O/bench_nbskip$s.err; done
timeout 600 compute-sanitizer --tool racecheck --print-limit 30 python tools/profile_step.py --case 12 --steps 6 --warmup 5 > $O/racecheck_sp.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 30 python tools/profile_step.py --case 12 --precision dp --steps 6 --warmup 5 > $O/racecheck_dp.log 2>&1
timeout 300 python bench.py --precision dp --case 48 --no-cpu-baseline --steps 300 --warmup 50 > $O/bench_dp48.json 2> $O/bench_dp48.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_dpd' -c 2 \
    -o $O/prof_dp python tools/profile_step.py --case 48 --precision dp --steps 5 > $O/prof_dp.log 2>&1
timeout 600 python bench.py --case 200 --no-cpu-baseline --steps 20 --warmup 5 > $O/bench_200.json 2> $O/bench_200.err
tail -3 $O/nbskip_pytest.log; tail -5 $O/racecheck_sp.log; tail -5 $O/racecheck_dp.log; cat $O/bench_nbskip0.json $O/bench_nbskip1.json $O/bench_dp48.json $O/bench_200.json | cut -c1-600
