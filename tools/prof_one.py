"""Dev tool: one pass of load + kmeans + counts + quantize (short, for ncu)."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qvz_b200 import lib
from qvz_b200.synth import synth_rows
from tests.helpers import synthetic_tables
ap = argparse.ArgumentParser()
ap.add_argument("--lines", type=int, default=1_000_000)
ap.add_argument("--columns", type=int, default=150)
ap.add_argument("--clusters", type=int, default=1)
ap.add_argument("--passes", type=int, default=1)
a = ap.parse_args()
n, c, k = a.lines, a.columns, a.clusters
rows = synth_rows(n, c, seed=1, device="cuda").cpu().numpy()
h = lib.Handle(0)
h.load_rows(rows, n, c, c + 1)
init = rows[[(i * 7919 + 13) % n for i in range(k)], :c]
t = synthetic_tables(k, c, seed=3)
seed = np.full(32, 0x55555555, np.uint32)
for _ in range(a.passes):
    r = h.kmeans(init, 4.0, want_ids=False)
    h.cond_counts(want=False)
    h.quantize(t, seed, want_symbols=False)
print("ok", r["iters"], h.timings())
