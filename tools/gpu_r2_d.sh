#!/bin/bash
# round 2, run D (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/d_tests.log 2>&1
rc=$?
tail -3 gpurun_out/d_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python bench.py --config cfg4 --lines 24000000 --steps 3 --warmup 2 --no-cpu > gpurun_out/d_cfg4_24M.json 2> gpurun_out/d_cfg4_24M.err
QVZ_KM_SMEM_MEANS=1 timeout 600 python bench.py --config cfg4 --lines 24000000 --steps 3 --warmup 2 --no-cpu --no-parity > gpurun_out/d_cfg4_24M_smem.json 2> gpurun_out/d_cfg4_24M_smem.err
timeout 900 python bench.py --steps 3 --warmup 2 --e2e-steps 2 --no-cpu > gpurun_out/d_cfg4_full.json 2> gpurun_out/d_cfg4_full.err
timeout 600 python bench.py --config cfg2 --steps 3 --warmup 2 --no-cpu --no-parity > gpurun_out/d_cfg2.json 2> gpurun_out/d_cfg2.err
