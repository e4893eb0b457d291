#!/bin/bash
# round 2, run R: walk with 1/C as a kernel parameter (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "quantize or golden or compiled_reference" > gpurun_out/r_tests.log 2>&1
tail -3 gpurun_out/r_tests.log
timeout 900 python bench.py --steps 3 --warmup 2 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/r_cfg4_full.json 2> gpurun_out/r_cfg4_full.err
timeout 600 python bench.py --config cfg2 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/r_cfg2.json 2> gpurun_out/r_cfg2.err
