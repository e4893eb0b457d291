set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t_final.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err || exit 1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qvz_ -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'batched|draws|planes' -c 3 -o gpurun_out/prof_r1h -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/t_final.log; python tools/brief.py final < gpurun_out/bench_native.json
