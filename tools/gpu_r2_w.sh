#!/bin/bash
# round 2, run W: one row per thread shapes of the k-means MMA kernel (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans_kernel_variants" > gpurun_out/w_tests.log 2>&1
rc=$?
tail -3 gpurun_out/w_tests.log
if [ $rc -ne 0 ]; then tail -50 gpurun_out/w_tests.log; exit 0; fi
B="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity"
for s in 1256 1128; do
QVZ_KM_SHAPE=$s QVZ_DEBUG_KM=1 timeout 600 $B > gpurun_out/w_cfg4_full_$s.json 2> gpurun_out/w_cfg4_full_$s.err
done
