timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench8.json 2> gpurun_out/bench8.err
python tools/brief.py n8 < gpurun_out/bench8.json || tail -20 gpurun_out/bench8.err
