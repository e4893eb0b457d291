#!/bin/bash
# round 2, run M: CTA-pooled MMA batches in k-means, walk without cluster code (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/m_tests.log 2>&1
rc=$?
tail -3 gpurun_out/m_tests.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/m_tests.log; exit 0; fi
QVZ_DEBUG_KM=1 QVZ_DEBUG_WALK=1 timeout 900 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/m_cfg4_full.json 2> gpurun_out/m_cfg4_full.err
QVZ_KM_SHAPE=464 QVZ_DEBUG_KM=1 timeout 900 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/m_cfg4_full_464.json 2> gpurun_out/m_cfg4_full_464.err
QVZ_KM_SHAPE=2256 QVZ_DEBUG_KM=1 timeout 900 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/m_cfg4_full_2256.json 2> gpurun_out/m_cfg4_full_2256.err
ls -la gpurun_out/m_*
