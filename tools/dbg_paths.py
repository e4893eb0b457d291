import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from tests.conftest import load_golden
from oracle.bindings import DEBUG_SEED
from qvz_b200 import lib
h = lib.Handle(0)
for name in ["small_f05_M_c2", "small_r2_L_c1"]:
    g = load_golden(name)
    c = g["columns"]
    h.load_rows(g["rows"], g["rows"].shape[0], c, g["rows"].shape[1])
    init = g["rows"][g["picks"].astype(np.int64), :c]
    r = h.kmeans(init, float(g["threshold"]))
    print(name, "kmeans ids ok", np.array_equal(r["ids"], g["ids"]), "counts ok", np.array_equal(h.cond_counts(), g["cond_counts"]))
    for rep in range(2):
        q = h.quantize(g["tables"], DEBUG_SEED, want_qv=True, want_err=True)
        os.environ["QVZ_FORCE_LINE_MAJOR"] = "1"
        q2 = h.quantize(g["tables"], DEBUG_SEED, want_qv=True, want_err=True)
        del os.environ["QVZ_FORCE_LINE_MAJOR"]
        for key in ("symbols", "qv", "line_err"):
            a, b = np.array_equal(q[key], g[key]), np.array_equal(q2[key], g[key])
            print("  ", rep, key, "batched ok", a, "line-major ok", b)
            if not a:
                bad = np.argwhere(q[key] != g[key])
                print("     batched first bad", bad[:5], "count", len(bad))
            if not b:
                bad = np.argwhere(q2[key] != g[key])
                print("     line-major first bad", bad[:5], "count", len(bad))
