#!/bin/bash
# round 2, pass U (8 GPUs): final multi-GPU check of the round's last library -- 8-rank parity incl. the channel, the gang, the decks, the bench
O=gpurun_out/r2u; mkdir -p $O
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "test_multi_gpu_parity and (8-L12 or channel)" > $O/pytest_mgpu.log 2>&1; echo "exit $?" >> $O/pytest_mgpu.log
tail -4 $O/pytest_mgpu.log
timeout 300 python -m pytest tests/test_gang.py -m gpu -q > $O/pytest_gang.log 2>&1; echo "exit $?" >> $O/pytest_gang.log
tail -3 $O/pytest_gang.log
timeout 300 python -m pytest tests/test_lammps_deck.py -m gpu -q -k "several_gpus" > $O/pytest_deck.log 2>&1; echo "exit $?" >> $O/pytest_deck.log
tail -3 $O/pytest_deck.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29731 bench.py --gpus 8 --steps 100 --warmup 10 > $O/bench_200_n8.json 2> $O/bench_200_n8.err
python - <<PY
import json
for f in ('bench_200_n8',):
    try:
        d=json.load(open('$O/%s.json'%f))
        print(f, '%.3e'%d['value'], 'e2e', d['e2e'] and '%.3e'%d['e2e']['value'], d['ms_per_step'], {k:(round(v['ms_total']/max(v['calls'],1),4),v['calls']) for k,v in d['phases'].items()}, d.get('parity_check',{}).get('ok'))
    except Exception as e:
        print(f, 'failed', e); print(open('$O/%s.err'%f).read()[-1500:])
PY
