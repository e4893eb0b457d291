#!/bin/bash
# round 2, run F: register-blocked incremental k-means kernel (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans or golden or compiled_reference or sharded or cfg_scale" > gpurun_out/f_tests.log 2>&1
rc=$?
tail -3 gpurun_out/f_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
B="python bench.py --config cfg4 --lines 24000000 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1"
timeout 600 $B > gpurun_out/f_cfg4_24M.json 2> gpurun_out/f_cfg4_24M.err
QVZ_KM_TILED_INCR=1 timeout 600 $B > gpurun_out/f_cfg4_24M_tiled.json 2> gpurun_out/f_cfg4_24M_tiled.err
for n in 2 3 6; do QVZ_KM_CTAS=$n timeout 600 $B > gpurun_out/f_cfg4_24M_ctas$n.json 2> gpurun_out/f_cfg4_24M_ctas$n.err; done
timeout 900 python bench.py --steps 3 --warmup 2 --e2e-steps 1 --no-cpu > gpurun_out/f_cfg4_full.json 2> gpurun_out/f_cfg4_full.err
CMD="python bench.py --config cfg4 --lines 12000000 --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 1"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qvz_kmeans_assign_incr4 -s 5 -c 2 -o gpurun_out/f_assign $CMD > gpurun_out/f_ncu_assign.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qvz_quantize_batched -s 1 -c 1 -o gpurun_out/f_walk $CMD > gpurun_out/f_ncu_walk.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qvz_draws_kernel -s 1 -c 1 -o gpurun_out/f_draws $CMD > gpurun_out/f_ncu_draws.log 2>&1
ls -la gpurun_out/f_*
