#!/bin/bash
# round 2, run O: whole GPU suite, default bench (timed), all configs (dev script)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/o_tests.log 2>&1
tail -5 gpurun_out/o_tests.log
( time python bench.py ) > gpurun_out/o_bench_default.json 2> gpurun_out/o_bench_default.err
tail -4 gpurun_out/o_bench_default.err
( time python bench.py --impl reference ) > gpurun_out/o_bench_reference.json 2> gpurun_out/o_bench_reference.err
tail -4 gpurun_out/o_bench_reference.err
for c in cfg2 cfg3 cfg5; do
QVZ_DEBUG_WALK=1 timeout 900 python bench.py --config $c --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/o_$c.json 2> gpurun_out/o_$c.err
done
ls -la gpurun_out/o_*
