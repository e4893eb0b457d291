#!/bin/bash
# round 2, run K: cluster-multicast table staging in the walk (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "quantize or golden or compiled_reference or sharded or cfg_scale or prefetched" > gpurun_out/k_tests.log 2>&1
rc=$?
tail -3 gpurun_out/k_tests.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/k_tests.log; exit 0; fi
for cs in 4 2 1; do
QVZ_WALK_CLUSTER=$cs QVZ_DEBUG_WALK=1 timeout 600 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/k_cfg4_full_cs$cs.json 2> gpurun_out/k_cfg4_full_cs$cs.err
QVZ_WALK_CLUSTER=$cs QVZ_DEBUG_WALK=1 timeout 600 python bench.py --config cfg2 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/k_cfg2_cs$cs.json 2> gpurun_out/k_cfg2_cs$cs.err
done
QVZ_DEBUG_WALK=1 timeout 600 python bench.py --config cfg3 --lines 24000000 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/k_cfg3_24M.json 2> gpurun_out/k_cfg3_24M.err
QVZ_DEBUG_WALK=1 timeout 600 python bench.py --config cfg5 --lines 16000000 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/k_cfg5_16M.json 2> gpurun_out/k_cfg5_16M.err
CMD="python bench.py --config cfg4 --lines 12000000 --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 1"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qvz_quantize_batched -s 1 -c 1 -o gpurun_out/k_walk $CMD > gpurun_out/k_ncu_walk.log 2>&1
ls -la gpurun_out/k_*
