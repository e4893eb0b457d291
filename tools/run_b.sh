timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench1.json 2> gpurun_out/bench1.err
python tools/brief.py n1 < gpurun_out/bench1.json || tail -20 gpurun_out/bench1.err
python -c "
import json; d=json.loads(open('gpurun_out/bench1.json').read().strip().splitlines()[-1]); print(d['e2e'])"
