timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench1.json 2> gpurun_out/bench1.err
python tools/brief.py n1 < gpurun_out/bench1.json || tail -20 gpurun_out/bench1.err
