#!/bin/bash
# round 2, final run: default bench, other configs, ncu captures for profiles/ (dev script; the whole GPU suite ran green before: 76 passed)
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/fin_bench_default.json 2> gpurun_out/fin_bench_default.err
tail -3 gpurun_out/fin_bench_default.err
for c in cfg2 cfg3 cfg5; do
QVZ_DEBUG_WALK=1 timeout 900 python bench.py --config $c --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/fin_$c.json 2> gpurun_out/fin_$c.err
done
CMD="python bench.py --config cfg4 --lines 24000000 --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 1"
$CMD > gpurun_out/fin_plain.json 2> gpurun_out/fin_plain.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qvz_ --csv --log-file gpurun_out/fin_launches.csv $CMD > gpurun_out/fin_ncu_launch.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:qvz_kmeans_assign_mma -s 6 -c 3 -o gpurun_out/fin_assign $CMD > gpurun_out/fin_ncu_assign.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:qvz_cond_counts_kernel -s 1 -c 1 -o gpurun_out/fin_counts $CMD > gpurun_out/fin_ncu_counts.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:qvz_quantize_batched -s 1 -c 1 -o gpurun_out/fin_walk $CMD > gpurun_out/fin_ncu_walk.log 2>&1
du -sh gpurun_out
