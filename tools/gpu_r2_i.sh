#!/bin/bash
# round 2, run I: tile shapes of the k-means MMA kernel (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans or golden or compiled_reference or sharded or cfg_scale" > gpurun_out/i_tests.log 2>&1
rc=$?
tail -3 gpurun_out/i_tests.log
if [ $rc -ne 0 ]; then tail -40 gpurun_out/i_tests.log; exit 0; fi
for s in 2128 264 2256 464; do
QVZ_KM_SHAPE=$s QVZ_DEBUG_KM=1 timeout 900 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/i_cfg4_full_$s.json 2> gpurun_out/i_cfg4_full_$s.err
done
ls -la gpurun_out/i_*
