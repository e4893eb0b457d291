#!/bin/bash
# round 2, run B: ncu launch list + full captures of the three main kernels on a cfg4-shaped shard (dev script)
mkdir -p gpurun_out
CMD="python bench.py --config cfg4 --lines 24000000 --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 1"
$CMD > gpurun_out/b_plain.json 2> gpurun_out/b_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches.csv $CMD > gpurun_out/b_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qvz_quantize_batched -s 1 -c 1 -o gpurun_out/b_walk $CMD > gpurun_out/b_ncu_walk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qvz_cond_counts_kernel -s 1 -c 1 -o gpurun_out/b_counts $CMD > gpurun_out/b_ncu_counts.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qvz_kmeans_assign -s 7 -c 2 -o gpurun_out/b_assign $CMD > gpurun_out/b_ncu_assign.log 2>&1
ls -la gpurun_out/b_*
