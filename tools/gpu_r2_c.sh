#!/bin/bash
# round 2, run C: tests, cfg4 benches, filtered ncu captures (dev script)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_tests.log 2>&1
rc=$?
tail -3 gpurun_out/c_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
QVZ_DEBUG_WALK=1 timeout 600 python bench.py --config cfg4 --lines 24000000 --steps 3 --warmup 2 --no-cpu > gpurun_out/c_cfg4_24M.json 2> gpurun_out/c_cfg4_24M.err
QVZ_DEBUG_WALK=1 timeout 900 python bench.py --steps 3 --warmup 2 --e2e-steps 2 --no-cpu > gpurun_out/c_cfg4_full.json 2> gpurun_out/c_cfg4_full.err
timeout 600 python bench.py --config cfg3 --lines 24000000 --steps 3 --warmup 2 --no-cpu --no-parity > gpurun_out/c_cfg3_24M.json 2> gpurun_out/c_cfg3_24M.err
timeout 600 python bench.py --config cfg5 --lines 16000000 --steps 3 --warmup 2 --no-cpu --no-parity > gpurun_out/c_cfg5_16M.json 2> gpurun_out/c_cfg5_16M.err
CMD="python bench.py --config cfg4 --lines 12000000 --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 1"
$CMD > gpurun_out/c_plain.json 2> gpurun_out/c_plain.err &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qvz_cond_counts_kernel -s 1 -c 1 -o gpurun_out/c_counts $CMD > gpurun_out/c_ncu_counts.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qvz_kmeans_assign -s 7 -c 2 -o gpurun_out/c_assign $CMD > gpurun_out/c_ncu_assign.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qvz_quantize_batched -s 1 -c 1 -o gpurun_out/c_walk $CMD > gpurun_out/c_ncu_walk.log 2>&1
ls -la gpurun_out/c_*
