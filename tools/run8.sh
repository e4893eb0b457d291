timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err
python tools/brief.py n4 < gpurun_out/bench4.json || tail -20 gpurun_out/bench4.err
