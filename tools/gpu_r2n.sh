#!/bin/bash
# round 2, pass N (1 GPU): warps per CTA of the tile build (segment barrier imbalance)
O=gpurun_out/r2n; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
for w in 5 6 7 8 9 10 12; do
  MESO_NB_WARPS=$w timeout 300 python bench.py --case 64 --no-cpu-baseline --no-e2e --no-parity --steps 200 --warmup 20 > $O/bench_w$w.json 2> $O/bench_w$w.err
  python - <<PY
import json
d=json.load(open('$O/bench_w$w.json'))
print('warps $w', '%.3e'%d['value'], {k:round(v['ms_total']/max(v['calls'],1),4) for k,v in d['phases'].items()})
PY
done
