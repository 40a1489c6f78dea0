#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fixes.py -m gpu -x -q > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "one_gpu and (2-L12 or polymer or 4-L8)" > $O/pytest_mr.log 2>&1; echo "exit $?" >> $O/pytest_mr.log
tail -3 $O/pytest_mr.log
for r in 0 1; do
MESO_SORT_RADIX=$r timeout 300 python bench.py --case 64 --no-cpu-baseline --no-e2e --no-parity --steps 300 --warmup 50 > $O/bench_case64_r$r.json 2> $O/bench_case64_r$r.err
python -c "
import json
d=json.load(open('$O/bench_case64_r$r.json'))
print('radix=$r case64', '%.3e'%d['value'], {k:(round(v['ms_total']/max(v['calls'],1),4),v['calls']) for k,v in d['phases'].items()})"
done
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 50 --warmup 5 > $O/bench_200.json 2> $O/bench_200.err
python -c "
import json
d=json.load(open('$O/bench_200.json'))
print('200^3', '%.3e'%d['value'], {k:(round(v['ms_total']/max(v['calls'],1),4),v['calls']) for k,v in d['phases'].items()})"
