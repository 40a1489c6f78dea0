#!/bin/bash
# round 2, pass K (8 GPUs): gang handle with one brick per GPU (2, 4, 8 bricks; fluid, chains, phase API), the deck through lmp_meso_b200
# on 2 / 4 / 8 GPUs, and sp.run case 128 on 8 GPUs through the LAMMPS binary
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m pytest tests/test_gang.py -m gpu -q > $O/pytest_gang.log 2>&1; echo "exit $?" >> $O/pytest_gang.log
tail -6 $O/pytest_gang.log
timeout 900 python -m pytest tests/test_gang.py -m gpu -q > $O/pytest_gang2.log 2>&1; echo "exit $?" >> $O/pytest_gang2.log
tail -3 $O/pytest_gang2.log
timeout 900 python -m pytest tests/test_lammps_deck.py -m gpu -q -k "several_gpus or thermo_only" > $O/pytest_deck.log 2>&1; echo "exit $?" >> $O/pytest_deck.log
tail -6 $O/pytest_deck.log
python - <<PY > $O/deck128.log 2>&1
import os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
from meso_b200 import workload
sys.path.insert(0, 'tests')
import test_lammps_deck as T
d = '/tmp/deck128'; os.makedirs(d, exist_ok=True)
L = 128
workload.write_data(d + '/%d.data' % L, workload.dpd_fluid(L), L)
open(d + '/in.run', 'w').write(T.DECK.format(prec='sp', pair='dpd/fast/meso', extra='', thermo=100, steps=1000, dump=''))
for devs in ('0-7',):
    t0 = time.time()
    out = subprocess.run([T.LMP, '-in', 'in.run', '-var', 'case', str(L), '-log', 'none'], cwd=d, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, MESO_DEVICES=devs))
    print('MESO_DEVICES', devs, 'rc', out.returncode, 'wall %.1f s' % (time.time() - t0))
    print('\n'.join(s for s in out.stdout.split('\n') if 'Loop time' in s or 'bricks' in s or s.startswith('Pair') or s.startswith('Neigh') or s.startswith('Comm') or 'ERROR' in s))
    for s in out.stdout.split('\n'):
        if s.startswith('Loop time'):
            t = float(s.split()[3]); print('particle-steps/s %.3e' % (4 * L ** 3 * 1000 / t))
    print(out.stderr[-500:])
PY
cat $O/deck128.log
