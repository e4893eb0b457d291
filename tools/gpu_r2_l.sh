#!/bin/bash
# round 2, run L: cluster-multicast A/B (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "quantize or golden or compiled_reference" > gpurun_out/l_tests.log 2>&1
rc=$?
tail -3 gpurun_out/l_tests.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/l_tests.log; exit 0; fi
for cs in 1 2 4; do
QVZ_WALK_CLUSTER=$cs QVZ_DEBUG_WALK=1 timeout 600 python bench.py --lines 48000000 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/l_cfg4_48M_cs$cs.json 2> gpurun_out/l_cfg4_48M_cs$cs.err
QVZ_WALK_CLUSTER=$cs QVZ_DEBUG_WALK=1 timeout 600 python bench.py --config cfg2 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/l_cfg2_cs$cs.json 2> gpurun_out/l_cfg2_cs$cs.err
done
ls -la gpurun_out/l_*
