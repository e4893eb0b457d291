timeout 600 python tools/prof_one.py --lines 4000000 --clusters 5 > gpurun_out/plain5.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'assign|cond_counts|batched|draws' -c 16 -o gpurun_out/prof_r1_k5c -f python tools/prof_one.py --lines 4000000 --clusters 5 > gpurun_out/ncu5.log 2>&1
tail -2 gpurun_out/ncu5.log
