#!/bin/bash
# round 2, run X: scalar path for warps with few changed rows (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans or golden" > gpurun_out/y_tests.log 2>&1
rc=$?
tail -3 gpurun_out/y_tests.log
if [ $rc -ne 0 ]; then tail -50 gpurun_out/y_tests.log; exit 0; fi
B="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity"
QVZ_DEBUG_KM=1 timeout 600 $B > gpurun_out/y_cfg4_full.json 2> gpurun_out/y_cfg4_full.err
