"""Dev tool: per-stage device timings of the CUDA path on synthetic data (not the bench contract)."""
import argparse
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qvz_b200 import lib
from qvz_b200.synth import synth_rows
from tests.helpers import synthetic_tables

ap = argparse.ArgumentParser()
ap.add_argument("--lines", type=int, default=4_000_000)
ap.add_argument("--columns", type=int, default=150)
ap.add_argument("--clusters", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tables", default="")
a = ap.parse_args()

n, c, k = a.lines, a.columns, a.clusters
t0 = time.time()
rows_d = synth_rows(n, c, seed=1, device="cuda")
torch.cuda.synchronize()
rows = torch.empty(rows_d.shape, dtype=torch.uint8, pin_memory=True)
rows.copy_(rows_d)
del rows_d
torch.cuda.empty_cache()
print(f"synth {time.time()-t0:.2f}s  rows {tuple(rows.shape)}")
h = lib.Handle(0)
h.load_rows(rows, n, c, c + 1)
rn = rows.numpy()
picks = [(i * 7919 + 13) % n for i in range(k)]
init = rn[picks, :c]
if a.tables:
    from oracle.bindings import FlatTables
    tables = FlatTables.load(a.tables)
else:
    tables = synthetic_tables(k, c, seed=3)
seed = np.full(32, 0x55555555, np.uint32)
sym = torch.empty((n, c), dtype=torch.uint8, pin_memory=True)
for rep in range(a.reps):
    t0 = time.time(); h.load_rows(rows, n, c, c + 1); tl = time.time() - t0
    t0 = time.time(); r = h.kmeans(init, 4.0, want_ids=False); tk = time.time() - t0
    t0 = time.time(); h.cond_counts(want=False); tc = time.time() - t0
    t0 = time.time(); h.quantize(tables, seed, want_symbols=False); tq = time.time() - t0
    t0 = time.time(); h.quantize(tables, seed, symbols_out=sym); tq2 = time.time() - t0
    tm = h.timings()
    sym_n = n * c
    print(f"rep{rep}: iters={r['iters']} counts={r['counts']}")
    print(f"  wall: load {tl*1e3:.1f} kmeans {tk*1e3:.1f} counts {tc*1e3:.1f} quant(resident) {tq*1e3:.1f} quant(+d2h) {tq2*1e3:.1f} ms")
    print("  dev : " + " ".join(f"{k_}={v:.3f}" if isinstance(v, float) else f"{k_}={v}" for k_, v in tm.items()))
    it = max(tm['kmeans_iters'], 1)
    print(f"  GB/s: kmeans-assign/iter {sym_n/(tm['kmeans_assign_ms']/it)/1e6:.0f}  counts {sym_n/tm['cond_counts_ms']/1e6:.0f}  "
          f"quantize {2*sym_n/tm['quantize_ms']/1e6:.0f}  ingest {2*sym_n/tm['load_layout_ms']/1e6:.0f}  h2d {sym_n/tm['load_h2d_ms']/1e6:.0f}")
