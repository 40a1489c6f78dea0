#!/bin/bash
# One GPU-box pass: parity tests, bench, launch list, full ncu capture of the top kernels.  Usage: tools/gpu_round.sh <tag>
TAG=${1:-run}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 600 python bench.py --steps 1000 --warmup 100 > $O/bench_sp.json 2> $O/bench_sp.err
timeout 600 python bench.py --steps 500 --warmup 50 --precision dp --case 48 --no-cpu-baseline > $O/bench_dp.json 2> $O/bench_dp.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_sp.csv \
    python tools/profile_step.py --case 64 --precision sp --steps 10 > $O/launches_sp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_dpd|k_build_neighbors' -c 4 \
    -o $O/prof_sp python tools/profile_step.py --case 64 --precision sp --steps 5 > $O/prof_sp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_dpd' -c 2 \
    -o $O/prof_dp python tools/profile_step.py --case 48 --precision dp --steps 5 > $O/prof_dp.log 2>&1
tail -3 $O/pytest.log; cat $O/bench_sp.json; tail -2 $O/bench_sp.err
