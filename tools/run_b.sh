timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench1.json 2> gpurun_out/bench1.err
python tools/brief.py early < gpurun_out/bench1.json || tail -20 gpurun_out/bench1.err
QVZ_NO_EARLY_DRAWS=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench1b.json 2> gpurun_out/bench1b.err
python tools/brief.py noearly < gpurun_out/bench1b.json || tail -20 gpurun_out/bench1b.err
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
