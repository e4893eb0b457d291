#!/bin/bash
# round 2, run S: ring of table-image buffers in the walk (dev script)
mkdir -p gpurun_out
QVZ_SKIP_FULL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "quantize or golden or compiled_reference or sharded" > gpurun_out/s_tests.log 2>&1
rc=$?
tail -3 gpurun_out/s_tests.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/s_tests.log; exit 0; fi
for s in 4 2 1; do
QVZ_WALK_S=$s QVZ_DEBUG_WALK=1 timeout 600 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/s_cfg4_full_S$s.json 2> gpurun_out/s_cfg4_full_S$s.err
done
for s in 4 2; do
QVZ_WALK_S=$s QVZ_DEBUG_WALK=1 timeout 600 python bench.py --config cfg2 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/s_cfg2_S$s.json 2> gpurun_out/s_cfg2_S$s.err
done
QVZ_WALK_NBUF=2 QVZ_DEBUG_WALK=1 timeout 600 python bench.py --config cfg2 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/s_cfg2_S4_nbuf2.json 2> gpurun_out/s_cfg2_S4_nbuf2.err
