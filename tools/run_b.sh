timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'batched|draws|planes' -c 6 -o gpurun_out/prof_r1h -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
