#!/bin/bash
# round 2, run Q: new tests, bench after the division change, ncu captures for profiles/ (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/q_tests.log 2>&1
rc=$?
tail -3 gpurun_out/q_tests.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/q_tests.log; exit 0; fi
timeout 900 python bench.py --steps 3 --warmup 2 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/q_cfg4_full.json 2> gpurun_out/q_cfg4_full.err
# launch list + full captures on a 24 M-line shard of cfg4 (same columns / clusters / tables; ncu replays every kernel ~40 times)
CMD="python bench.py --config cfg4 --lines 24000000 --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 1"
$CMD > gpurun_out/q_plain.json 2> gpurun_out/q_plain.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qvz_ --csv --log-file gpurun_out/q_launches.csv $CMD > gpurun_out/q_ncu_launch.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:qvz_kmeans_assign_mma -s 6 -c 6 -o gpurun_out/q_assign $CMD > gpurun_out/q_ncu_assign.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:qvz_cond_counts_kernel -s 1 -c 1 -o gpurun_out/q_counts $CMD > gpurun_out/q_ncu_counts.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:qvz_quantize_batched -s 1 -c 1 -o gpurun_out/q_walk $CMD > gpurun_out/q_ncu_walk.log 2>&1
ls -la gpurun_out/q_*
