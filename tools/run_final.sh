set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t_final.log
timeout 900 python bench.py > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err
cat gpurun_out/t_final.log; python tools/brief.py final < gpurun_out/bench_native.json || tail gpurun_out/bench_native.err
