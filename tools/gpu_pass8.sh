#!/bin/bash
# 2-GPU pass: parity vs the simulated-rank oracle, deck tests, weak-scaling bench at N = 2
TAG=${1:-pass8}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_lammps_deck.py tests/test_gpu_parity.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -6 $O/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 400 --warmup 50 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 200 --warmup 20 --case 100 > $O/bench_n2_c100.json 2> $O/bench_n2_c100.err
timeout 600 python bench.py --gpus 1 --steps 400 --warmup 50 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
python - <<PY
import json
for f in ("bench_n1","bench_n2","bench_n2_c100"):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], "roof %.3f"%d["roofline"]["frac"], {k:(round(v["ms_total"]/max(v["calls"],1)*1e3,1)) for k,v in d["phases"].items()})
    except Exception as e: print(f, "FAILED", e, open("$O/%s.err"%f).read()[-800:])
PY
