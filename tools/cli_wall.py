"""Dev tool: wall clock of the whole command line on cfg1 (1M x 100, -q -f 1.0 -d M -c 1), reference binary vs qvz_b200/host/qvz."""
import os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qvz_b200.synth import synth_rows
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF, CLI = os.path.join(ROOT, "oracle", "_ref", "qvz_ref_det"), os.path.join(ROOT, "qvz_b200", "host", "qvz")
n, c = 1_000_000, 100
rows = synth_rows(n, c, seed=1234, device="cuda").cpu().numpy()
with tempfile.TemporaryDirectory() as d:
    src = d + "/cfg1.txt"
    rows.tofile(src)
    res = {}
    for tag, exe, env in (("new", CLI, {"QVZ_DEBUG_SEED": "1"}), ("new_again", CLI, {"QVZ_DEBUG_SEED": "1"}), ("ref", REF, {})):
        t0 = time.perf_counter()
        r = subprocess.run([exe, "-q", "-f", "1.0", "-d", "M", "-c", "1", "-s", src, d + f"/{tag}.qvz"], env={**os.environ, **env}, capture_output=True, text=True)
        res[tag] = time.perf_counter() - t0
        print(tag, f"{res[tag]:.2f} s wall;  -s line: {r.stdout.strip()[:200]}")
    same = np.array_equal(np.fromfile(d + "/new.qvz", np.uint8), np.fromfile(d + "/ref.qvz", np.uint8))
    print("identical files:", same, " speed-up of the whole command:", round(res["ref"] / res["new_again"], 1))
