"""One line per bench.py JSON output: value, step, stage times and per-kernel GB/s (dev helper)."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    r = d['roofline']; s = d['stage_ms']; st = d['steps']
    print(f"{f.split('/')[-1]:32s} val {d['value']:8.1f} ms {d['ms_per_step']:8.3f} it {d['config']['kmeans_iterations']:2d} km {s['kmeans_ms']/st:7.3f} asg {s['kmeans_assign_ms']/st:7.3f} cc {s['cond_counts_ms']/st:7.3f} q {s['quantize_ms']/st:7.3f} dr {s['quantize_draws_ms']/st:6.3f} set {s['quantize_setup_ms']/st:5.3f} whole {r['whole_step']['frac']:.3f} | " + ' '.join(f"{k.replace('qvz_','').replace('_kernel','')}={v:.0f}" for k, v in r['per_kernel_GBps'].items()))
