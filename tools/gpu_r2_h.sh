#!/bin/bash
# round 2, run H: k-means with tensor-core column sums (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans or golden or compiled_reference or sharded or cfg_scale" > gpurun_out/h_tests.log 2>&1
rc=$?
tail -3 gpurun_out/h_tests.log
if [ $rc -ne 0 ]; then tail -40 gpurun_out/h_tests.log; exit 0; fi
B="python bench.py --config cfg4 --lines 24000000 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1"
QVZ_DEBUG_KM=1 timeout 600 $B > gpurun_out/h_cfg4_24M.json 2> gpurun_out/h_cfg4_24M.err
for n in 32 128; do QVZ_KM_NT=$n timeout 600 $B > gpurun_out/h_cfg4_24M_nt$n.json 2> gpurun_out/h_cfg4_24M_nt$n.err; done
QVZ_DEBUG_KM=1 QVZ_KM_SORTED=1 timeout 600 $B > gpurun_out/h_cfg4_24M_sorted.json 2> gpurun_out/h_cfg4_24M_sorted.err
QVZ_DEBUG_KM=1 timeout 900 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu > gpurun_out/h_cfg4_full.json 2> gpurun_out/h_cfg4_full.err
QVZ_DEBUG_KM=1 QVZ_KM_SORTED=1 timeout 900 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/h_cfg4_full_sorted.json 2> gpurun_out/h_cfg4_full_sorted.err
CMD="python bench.py --config cfg4 --lines 12000000 --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 1"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qvz_kmeans_assign_mma -s 6 -c 3 -o gpurun_out/h_assign $CMD > gpurun_out/h_ncu_assign.log 2>&1
ls -la gpurun_out/h_*
