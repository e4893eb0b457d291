import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from qvz_b200 import lib, hostlib
from qvz_b200.synth import synth_rows
from tests.helpers import synthetic_tables
n, c = 20_000_000, 150
d = synth_rows(n, c, seed=1, device="cuda")
rows = torch.empty(d.shape, dtype=torch.uint8, pin_memory=True); rows.copy_(d); torch.cuda.synchronize(); del d
h = lib.Handle(0)
t = synthetic_tables(1, c, seed=3)
seed = np.full(32, 0x55555555, np.uint32)
sym = torch.empty((n, c), dtype=torch.uint8, pin_memory=True)
ids = torch.empty(n, dtype=torch.uint8, pin_memory=True)
init = rows.numpy()[[12345], :c]
for rep in range(3):
    t0 = time.perf_counter(); h.load_rows(rows, n, c, c + 1); t1 = time.perf_counter()
    h.kmeans(init, 4.0, ids_out=ids.numpy()); t2 = time.perf_counter()
    h.cond_counts(); t3 = time.perf_counter()
    h.quantize(t, seed, symbols_out=sym); t4 = time.perf_counter()
    tm = h.timings()
    print(f"load {1e3*(t1-t0):.1f} ms ({n*(c+1)/(t1-t0)/1e9:.1f} GB/s)  kmeans+ids {1e3*(t2-t1):.1f}  counts {1e3*(t3-t2):.1f}  quantize+d2h {1e3*(t4-t3):.1f} ms ({n*c/(t4-t3)/1e9:.1f} GB/s)", {k: round(v, 2) for k, v in tm.items() if 'ms' in k})
