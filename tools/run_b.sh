./tools/ubench/atoms > gpurun_out/atoms.log 2>&1
grep -E "lane = bank|32 distinct banks|random" gpurun_out/atoms.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
