#!/bin/bash
# round 2, run T: walk with 512 threads x 8 lines against 1024 x 4 (dev script)
mkdir -p gpurun_out
V=$PWD/qvz_b200/csrc/libqvz_gpu_v512.so
QVZ_GPU_LIB=$V QVZ_SKIP_FULL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "quantize or golden or compiled_reference" > gpurun_out/t_tests.log 2>&1
tail -3 gpurun_out/t_tests.log
timeout 600 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/t_cfg4_full_1024.json 2> gpurun_out/t_cfg4_full_1024.err
QVZ_GPU_LIB=$V timeout 600 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/t_cfg4_full_512.json 2> gpurun_out/t_cfg4_full_512.err
timeout 600 python bench.py --config cfg2 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/t_cfg2_1024.json 2> gpurun_out/t_cfg2_1024.err
QVZ_GPU_LIB=$V timeout 600 python bench.py --config cfg2 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/t_cfg2_512.json 2> gpurun_out/t_cfg2_512.err
QVZ_GPU_LIB=$V timeout 600 python bench.py --config cfg5 --lines 16000000 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/t_cfg5_512.json 2> gpurun_out/t_cfg5_512.err
timeout 600 python bench.py --config cfg5 --lines 16000000 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/t_cfg5_1024.json 2> gpurun_out/t_cfg5_1024.err
