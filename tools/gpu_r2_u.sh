#!/bin/bash
# round 2, run U: double-buffered k-means tiles (dev script)
mkdir -p gpurun_out
QVZ_KM_NBUF=2 QVZ_SKIP_FULL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans or golden" > gpurun_out/u_tests.log 2>&1
tail -3 gpurun_out/u_tests.log
B="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity"
QVZ_DEBUG_KM=1 timeout 600 $B > gpurun_out/u_cfg4_full.json 2> gpurun_out/u_cfg4_full.err
for s in 2128 264 2256 464; do
QVZ_KM_NBUF=2 QVZ_KM_SHAPE=$s QVZ_DEBUG_KM=1 timeout 600 $B > gpurun_out/u_cfg4_full_db$s.json 2> gpurun_out/u_cfg4_full_db$s.err
done
