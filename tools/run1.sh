set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t1.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -3 gpurun_out/t1.log; python tools/brief.py n1 < gpurun_out/bench1.json || tail -20 gpurun_out/bench1.err
