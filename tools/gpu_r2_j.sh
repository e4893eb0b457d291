#!/bin/bash
# round 2, run J: draw-free walk (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/j_tests.log 2>&1
rc=$?
tail -3 gpurun_out/j_tests.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/j_tests.log; exit 0; fi
QVZ_DEBUG_WALK=1 timeout 900 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu > gpurun_out/j_cfg4_full.json 2> gpurun_out/j_cfg4_full.err
QVZ_DEBUG_WALK=1 timeout 600 python bench.py --config cfg3 --lines 24000000 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/j_cfg3_24M.json 2> gpurun_out/j_cfg3_24M.err
QVZ_DEBUG_WALK=1 timeout 600 python bench.py --config cfg5 --lines 16000000 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/j_cfg5_16M.json 2> gpurun_out/j_cfg5_16M.err
QVZ_DEBUG_WALK=1 timeout 600 python bench.py --config cfg2 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/j_cfg2.json 2> gpurun_out/j_cfg2.err
CMD="python bench.py --config cfg4 --lines 12000000 --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 1"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qvz_quantize_batched -s 1 -c 1 -o gpurun_out/j_walk $CMD > gpurun_out/j_ncu_walk.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qvz_kmeans_assign_mma -s 9 -c 2 -o gpurun_out/j_assign $CMD > gpurun_out/j_ncu_assign.log 2>&1
ls -la gpurun_out/j_*
