#!/bin/bash
# round 2, run N: hybrid MMA phase, walk back to per-batch staging (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans or golden or compiled_reference or sharded or cfg_scale or quantize" > gpurun_out/n_tests.log 2>&1
rc=$?
tail -3 gpurun_out/n_tests.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/n_tests.log; exit 0; fi
QVZ_DEBUG_KM=1 QVZ_DEBUG_WALK=1 timeout 900 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/n_cfg4_full.json 2> gpurun_out/n_cfg4_full.err
ls -la gpurun_out/n_*
