#!/bin/bash
O=gpurun_out/r2r; mkdir -p $O
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -k "one_gpu and (channel or 4-L12)" > $O/pytest_channel.log 2>&1; echo "exit $?" >> $O/pytest_channel.log
tail -12 $O/pytest_channel.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python bench.py --workload polymer_channel --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_channel_n1.json 2> $O/bench_channel_n1.err
python -c "
import json
d=json.load(open('$O/bench_channel_n1.json'))
print('channel n1', '%.3e'%d['value'], {k:(round(v['ms_total']/max(v['calls'],1),4),v['calls']) for k,v in d['phases'].items()})"
