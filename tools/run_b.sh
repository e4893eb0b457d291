for cfg in cfg3 cfg4 cfg5; do
timeout 900 python bench.py --config $cfg --lines 8000000 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
python tools/brief.py $cfg < gpurun_out/bench_$cfg.json || tail -5 gpurun_out/bench_$cfg.err
done
