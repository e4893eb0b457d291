"""Host-side rows of SURVEY.md section 8f, timed next to the unmodified reference on the same input (CPU only).

    python tools/host_bench.py [--lines 300000] [--columns 100] > profiles/rNN_host_side.json

  f2  codebook design   qvz_host_design          vs generate_codebooks (through oracle/_ref/libqvzref.so)
  f1  symbol consumer   qvz_host_encode (coder)   vs the reference's whole encode() minus its k-means/stats/codebook time
  f4  decoder           qvz_host_decode           vs the reference's decode()
The reference's coder cannot be timed alone (it is interleaved with the quantize walk), so its figure includes the walk.
"""
import argparse, json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.bindings import DEBUG_SEED, FlatTables, Oracle, Ref, kmeans_init_lines
from qvz_b200 import hostlib
from qvz_b200.synth import synth_rows

ap = argparse.ArgumentParser()
ap.add_argument("--lines", type=int, default=300_000)
ap.add_argument("--columns", type=int, default=100)
a = ap.parse_args()
n, c, k = a.lines, a.columns, 1
mode, ratio, dist = hostlib.MODE_RATIO, 1.0, hostlib.DIST_MSE            # cfg1: -f 1.0 -d M -c 1
rows = synth_rows(n, c, seed=1234).numpy()
R, O = Ref(), Oracle()
s = R.session(rows, c, k, mode=mode, ratio=ratio, distortion=dist)
ids = s.kmeans(kmeans_init_lines(n, k, R.rand_stream(2 * k)))["ids"]
counts, _ = s.stats()
t0 = time.perf_counter(); t_ref = s.tables(); ref_design = time.perf_counter() - t0
t0 = time.perf_counter(); cb = hostlib.design_codebooks(counts, c, k, mode, ratio, dist); my_design = time.perf_counter() - t0
same_tables = all(np.array_equal(getattr(cb, f), getattr(t_ref, f)) for f in ("nctx", "ctx_of", "q_off", "qratio", "qmap", "smap"))
q = O.quantize(rows, c, ids, FlatTables(k, c, cb.nctx, cb.ctx_of, cb.q_off, cb.qratio, cb.qmap, cb.smap, cb.distortion), DEBUG_SEED, want_err=False)
with tempfile.TemporaryDirectory() as d:
    src, ref_out, my_out = d + "/in.txt", d + "/ref.qvz", d + "/my.qvz"
    rows.tofile(src)
    t0 = time.perf_counter(); R.encode_file(src, ref_out, None, clusters=k, mode=mode, ratio=ratio, distortion=dist); ref_encode = time.perf_counter() - t0
    t0 = time.perf_counter(); cb.encode(my_out, ids, q["symbols"], DEBUG_SEED); my_coder = time.perf_counter() - t0
    same_file = np.array_equal(np.fromfile(ref_out, np.uint8), np.fromfile(my_out, np.uint8))
    t0 = time.perf_counter(); R.decode_file(ref_out, d + "/ref.txt"); ref_decode = time.perf_counter() - t0
    t0 = time.perf_counter(); hostlib.decode_file(ref_out, d + "/my.txt"); my_decode = time.perf_counter() - t0
    same_decode = np.array_equal(np.fromfile(d + "/ref.txt", np.uint8), np.fromfile(d + "/my.txt", np.uint8))
sym = n * c
print(json.dumps({
    "input": f"{n} x {c} synthetic lines, -f 1.0 -d M -c 1 (cfg1 shape), host threads as available: the reference is single-threaded; qvz_host_design shares the contexts of a column over the threads, qvz_host_encode runs its model passes on worker threads (interval arithmetic on one), qvz_host_decode is sequential",
    "codebook_design_s": {"reference_generate_codebooks": round(ref_design, 2), "qvz_host_design": round(my_design, 2), "tables_identical": bool(same_tables)},
    "coder_Msym_per_s": {"reference_encode_minus_design_incl_its_walk": round(sym / max(ref_encode - ref_design, 1e-9) / 1e6, 2),
                         "qvz_host_encode": round(sym / my_coder / 1e6, 2), "file_identical": bool(same_file)},
    "decoder_Msym_per_s": {"reference_decode": round(sym / ref_decode / 1e6, 2), "qvz_host_decode": round(sym / my_decode / 1e6, 2),
                           "output_identical": bool(same_decode)}}, indent=1))
