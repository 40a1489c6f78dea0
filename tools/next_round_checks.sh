#!/bin/sh
# GPU checks that were left open when round 1 ran out of GPU minutes.  Run each line under gpurun.
# 1. direct halo routes at 8 ranks (verified at 2 and 4):
#      gpurun --gpus 8 --timeout 300 -- 'python -m pytest tests/test_multi_gpu.py -m gpu -x -q'
# 2. one-shot migration (written, never run): parity at 2 / 4 / 8 ranks, fluid and bead-spring chains
#      gpurun --gpus 2 --timeout 300 -- 'MESO_EXCH_ONESHOT=1 python -m pytest tests/test_multi_gpu.py -m gpu -x -q'
#      gpurun --gpus 8 --timeout 300 -- 'MESO_EXCH_ONESHOT=1 python -m pytest tests/test_multi_gpu.py -m gpu -x -q'
#    then A/B:  torchrun --nproc-per-node 8 bench.py --gpus 8   with and without MESO_EXCH_ONESHOT=1 (phases.rebuild)
# 3. stencil-cell skip in the neighbor build (written, never run): every bit-exact table test, then the build time
#      gpurun --timeout 300 -- 'MESO_NB_SKIP=1 python -m pytest tests/test_gpu_parity.py tests/test_fixes.py -m gpu -x -q; \
#                               for s in 0 1; do MESO_NB_SKIP=$s python bench.py --no-cpu-baseline --steps 300 --warmup 50; done'
# 4. racecheck of the warp-synchronous kernels:
#      gpurun --timeout 600 -- 'compute-sanitizer --tool racecheck python tools/profile_step.py --case 16 --steps 6 --warmup 5'
echo "see the comments in this file"
