#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
MESO_FORCE_COMM_PATH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_comm.csv \
    python tools/profile_step.py --case 64 --precision sp --steps 10 > $O/launches_comm.log 2>&1
python tools/launch_summary.py $O/launches_comm.csv 2>/dev/null | head -40
MESO_FORCE_COMM_PATH=1 timeout 300 python bench.py --case 64 --no-cpu-baseline --no-e2e --no-parity --steps 200 --warmup 20 > $O/bench_comm.json 2> $O/bench_comm.err
python -c "
import json
d=json.load(open('$O/bench_comm.json'))
print('comm path', '%.3e'%d['value'], {k:(round(v['ms_total']/max(v['calls'],1),4),v['calls']) for k,v in d['phases'].items()})"
