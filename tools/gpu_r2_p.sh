#!/bin/bash
# round 2, run P: 2 GPUs -- NCCL parity tests + strong-scaled bench (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 400 python -m pytest tests/test_dist_nccl_gpu.py -m gpu -x -q > gpurun_out/p_tests.log 2>&1
tail -3 gpurun_out/p_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/p_cfg4_n2.json 2> gpurun_out/p_cfg4_n2.err
echo "n2 rc=$?"
