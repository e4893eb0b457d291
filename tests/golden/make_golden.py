"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libqvzref.so).

Run in the build container only (needs /root/reference to have been compiled by oracle/Makefile):
    python tests/golden/make_golden.py
Each fixture holds the input rows and everything the reference computed from them through its own
functions: k-means (ids, means, counts, moved log, iterations), conditional counts, the flattened
cond_quantizer_list_t, the quantize walk (symbols, -u image, per-line error, distortion) under the
DEBUG WELL seed, and the bytes of the .qvz / -u files written by the reference's encode().
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.bindings import (DEBUG_SEED, DIST_LORENTZ, DIST_MANHATTAN, DIST_MSE, MODE_FIXED,  # noqa: E402
                             MODE_RATIO, Ref, kmeans_init_lines)
from qvz_b200.synth import synth_rows  # noqa: E402

CASES = {
    # name: lines, columns, clusters, threshold, mode, ratio, distortion, profile, seed
    "small_f05_M_c2": (3000, 16, 2, 4.0, MODE_RATIO, 0.5, DIST_MSE, "illumina", 11),
    "small_r2_L_c1": (2500, 22, 1, 4.0, MODE_FIXED, 2.0, DIST_LORENTZ, "illumina", 12),
    "small_f10_A_c3": (4000, 13, 3, 4.0, MODE_RATIO, 1.0, DIST_MANHATTAN, "miseq", 13),
}


def main():
    R = Ref()
    here = os.path.dirname(os.path.abspath(__file__))
    for name, (n, c, k, thr, mode, ratio, dist, profile, seed) in CASES.items():
        rows = synth_rows(n, c, seed=seed, profile=profile).numpy()
        picks = kmeans_init_lines(n, k, R.rand_stream(2 * k))
        s = R.session(rows, c, k, threshold=thr, mode=mode, ratio=ratio, distortion=dist)
        km = s.kmeans(picks)
        counts, totals = s.stats()
        t = s.tables()
        q = s.quantize(DEBUG_SEED)
        with tempfile.TemporaryDirectory() as d:
            src, dst, uf = os.path.join(d, "in.txt"), os.path.join(d, "out.qvz"), os.path.join(d, "u.txt")
            rows.tofile(src)
            R.encode_file(src, dst, uf, clusters=k, threshold=thr, mode=mode, ratio=ratio, distortion=dist)
            qvz = np.fromfile(dst, np.uint8)
            udump = np.fromfile(uf, np.uint8)
        assert np.array_equal(udump.reshape(n, c + 1), q["qv"]), "harness walk != reference encode() -u dump"
        np.savez_compressed(
            os.path.join(here, name + ".npz"), rows=rows, columns=c, clusters=k, threshold=thr, mode=mode,
            ratio=ratio, dist=dist, picks=np.array(picks, np.uint64), ids=km["ids"], iters=km["iters"],
            means=km["means"], kcounts=km["counts"], moved=km["moved"], cond_counts=counts,
            t_nctx=t.nctx, t_ctx_of=t.ctx_of, t_q_off=t.q_off, t_qratio=t.qratio, t_qmap=t.qmap,
            t_smap=t.smap, t_distortion=t.distortion, symbols=q["symbols"], qv=q["qv"],
            line_err=q["line_err"], distortion=q["distortion"], qvz=qvz)
        print(name, "iters", km["iters"], "counts", km["counts"], "qvz bytes", qvz.size,
              "file", os.path.getsize(os.path.join(here, name + ".npz")))


if __name__ == "__main__":
    main()
