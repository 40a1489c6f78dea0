#!/bin/bash
# One multi-GPU box pass (gpurun --gpus N -- 'bash tools/gpu_multi.sh <tag> [N]'): multi-rank parity with one process per GPU
# (fluid, chains, channel, phase API), the gang handle (one process for all GPUs), the decks through lmp_meso_b200 on 2/4/8 GPUs,
# sp.run case 128 on all GPUs through the LAMMPS binary, and the bench lines (200^3 box, 64^3 bricks, polymer channel).
TAG=${1:-multi}; N=${2:-8}
O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "test_multi_gpu_parity" > $O/pytest_mgpu.log 2>&1; echo "exit $?" >> $O/pytest_mgpu.log
tail -4 $O/pytest_mgpu.log
timeout 300 python -m pytest tests/test_gang.py -m gpu -q > $O/pytest_gang.log 2>&1; echo "exit $?" >> $O/pytest_gang.log
tail -3 $O/pytest_gang.log
timeout 300 python -m pytest tests/test_lammps_deck.py -m gpu -q -k "several_gpus" > $O/pytest_deck.log 2>&1; echo "exit $?" >> $O/pytest_deck.log
tail -3 $O/pytest_deck.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $N --master-port 29731 bench.py --gpus $N --steps 100 --warmup 10 > $O/bench_200_n$N.json 2> $O/bench_200_n$N.err
timeout 600 $TR --nproc-per-node $N --master-port 29732 bench.py --gpus $N --case 64 --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_case64_n$N.json 2> $O/bench_case64_n$N.err
timeout 600 $TR --nproc-per-node $N --master-port 29733 bench.py --gpus $N --workload polymer_channel --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_channel_n$N.json 2> $O/bench_channel_n$N.err
python - <<PY
import json
for f in ('bench_200_n$N', 'bench_case64_n$N', 'bench_channel_n$N'):
    try:
        d = json.load(open('$O/%s.json' % f))
        print(f, '%.3e' % d['value'], 'e2e', d['e2e'] and '%.3e' % d['e2e']['value'], d['ms_per_step'],
              {k: (round(v['ms_total'] / max(v['calls'], 1), 4), v['calls']) for k, v in d['phases'].items()}, d.get('parity_check', {}).get('ok'))
    except Exception as e:
        print(f, 'failed', e); print(open('$O/%s.err' % f).read()[-1500:])
PY
python - <<PY > $O/deck128.log 2>&1
import os, subprocess, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
from meso_b200 import workload
import test_lammps_deck as T
d = '/tmp/deck128'; os.makedirs(d, exist_ok=True)
L = 128
workload.write_data(d + '/%d.data' % L, workload.dpd_fluid(L), L)
open(d + '/in.run', 'w').write(T.DECK.format(prec='sp', pair='dpd/fast/meso', extra='', thermo=100, steps=1000, dump=''))
out = subprocess.run([T.LMP, '-in', 'in.run', '-var', 'case', str(L), '-log', 'none'], cwd=d, capture_output=True, text=True, timeout=900,
                     env=dict(os.environ, MESO_DEVICES='0-%d' % ($N - 1)))
print('MESO_DEVICES 0-%d' % ($N - 1), 'rc', out.returncode)
for s in out.stdout.split('\n'):
    if 'Loop time' in s or 'bricks' in s or s[:4] in ('Pair', 'Neig', 'Comm') or 'ERROR' in s:
        print(s)
    if s.startswith('Loop time'):
        print('particle-steps/s %.3e' % (4 * L ** 3 * 1000 / float(s.split()[3])))
print(out.stderr[-500:])
PY
cat $O/deck128.log
