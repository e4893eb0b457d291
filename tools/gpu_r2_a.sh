#!/bin/bash
# round 2, run A: environment probe, GPU parity tests, first cfg4 / cfg2 bench lines (dev script)
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2
cat /sys/fs/cgroup/memory.max 2>/dev/null
ulimit -l
} > gpurun_out/a_env.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_tests.log 2>&1
rc=$?
echo "tests rc=$rc" >> gpurun_out/a_tests.log
tail -5 gpurun_out/a_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python bench.py --config cfg4 --lines 24000000 --steps 3 --warmup 2 --no-cpu > gpurun_out/a_cfg4_24M.json 2> gpurun_out/a_cfg4_24M.err
QVZ_NO_SUPPORT=1 timeout 600 python bench.py --config cfg4 --lines 24000000 --steps 3 --warmup 2 --no-cpu --no-parity > gpurun_out/a_cfg4_24M_nosup.json 2> gpurun_out/a_cfg4_24M_nosup.err
QVZ_NO_REACH=1 timeout 600 python bench.py --config cfg4 --lines 24000000 --steps 3 --warmup 2 --no-cpu --no-parity > gpurun_out/a_cfg4_24M_noreach.json 2> gpurun_out/a_cfg4_24M_noreach.err
timeout 600 python bench.py --config cfg2 --steps 5 --warmup 3 --no-cpu --no-parity > gpurun_out/a_cfg2.json 2> gpurun_out/a_cfg2.err
timeout 900 python bench.py --steps 3 --warmup 2 --e2e-steps 2 > gpurun_out/a_cfg4_full.json 2> gpurun_out/a_cfg4_full.err
echo "full rc=$?" >> gpurun_out/a_cfg4_full.err
tail -3 gpurun_out/a_tests.log
