#!/bin/bash
# round 2, pass Q (4 GPUs): BASELINE configs[4] (amphiphilic chains in a driven channel) across ranks -- parity on 2 and 4 real GPUs
# (one process per GPU, and one process for all GPUs through the gang handle), the deck test on 2 / 4 GPUs, the channel bench line
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "test_multi_gpu_parity and channel" > $O/pytest_channel.log 2>&1; echo "exit $?" >> $O/pytest_channel.log
tail -4 $O/pytest_channel.log
timeout 900 python -m pytest tests/test_gang.py -m gpu -q -k "channel or 2-12" > $O/pytest_gang_channel.log 2>&1; echo "exit $?" >> $O/pytest_gang_channel.log
tail -4 $O/pytest_gang_channel.log
timeout 900 python -m pytest tests/test_lammps_deck.py -m gpu -q -k "several_gpus" > $O/pytest_deck.log 2>&1; echo "exit $?" >> $O/pytest_deck.log
tail -4 $O/pytest_deck.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 4 --master-port 29721 bench.py --gpus 4 --workload polymer_channel --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_channel_n4.json 2> $O/bench_channel_n4.err
timeout 600 python bench.py --workload polymer_channel --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_channel_n1.json 2> $O/bench_channel_n1.err
python - <<PY
import json
for f in ('bench_channel_n4','bench_channel_n1'):
    try:
        d=json.load(open('$O/%s.json'%f))
        print(f, '%.3e'%d['value'], 'e2e', d['e2e'] and '%.3e'%d['e2e']['value'], d['ms_per_step'], {k:(round(v['ms_total']/max(v['calls'],1),4),v['calls']) for k,v in d['phases'].items()}, d['config']['temperature_end'])
    except Exception as e:
        print(f, 'failed', e); print(open('$O/%s.err'%f).read()[-1500:])
PY
