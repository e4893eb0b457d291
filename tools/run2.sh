set -x
timeout 900 python -m pytest tests/test_dist_nccl_gpu.py "tests/test_cli_gpu.py::test_cli_vs_reference_binary" -x -q 2>&1 | tail -4 > gpurun_out/t2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
tail -3 gpurun_out/t2.log; python tools/brief.py n2 < gpurun_out/bench2.json || cat gpurun_out/bench2.json gpurun_out/bench2.err | tail -20
