#!/bin/bash
# round 2, run Z: threshold of the scalar path (dev script)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity"
for n in 0 1 2; do
QVZ_KM_SCALAR_ROWS=$n QVZ_DEBUG_KM=1 timeout 600 $B > gpurun_out/z_cfg4_full_$n.json 2> gpurun_out/z_cfg4_full_$n.err
done
