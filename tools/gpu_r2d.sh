#!/bin/bash
# round 2, pass D (1 GPU): neighbor-build clipping variants (MESO_NB_CLIP) + ncu of the build kernel
O=gpurun_out/r2d; mkdir -p $O
for c in 0 1 2; do
  MESO_NB_CLIP=$c timeout 300 python bench.py --case 64 --no-cpu-baseline --no-e2e --no-parity --steps 200 --warmup 20 > $O/bench_clip$c.json 2> $O/bench_clip$c.err
done
MESO_NB_CLIP=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "setup_parity or plain_walk or production_rows or dense or boundaries" > $O/pytest_clip1.log 2>&1; tail -2 $O/pytest_clip1.log
for c in 1 2; do
MESO_NB_CLIP=$c timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_build_rows$' -c 1 \
    -o $O/prof_build_clip$c python tools/profile_step.py --case 64 --precision sp --steps 10 > $O/prof_build_clip$c.log 2>&1
done
for c in 0 1 2; do python - <<PY
import json
d=json.load(open('$O/bench_clip$c.json'))
print('clip$c', '%.3e'%d['value'], {k:(round(v['ms_total']/max(v['calls'],1),4),v['calls']) for k,v in d['phases'].items()})
PY
done
