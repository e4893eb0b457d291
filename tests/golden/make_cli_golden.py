#!/usr/bin/env python
"""Reference hashes for the command-line parity tests at the configurations' real read lengths.

Runs the UNMODIFIED reference binary (oracle/_ref/qvz_ref_det: the reference compiled from /root/reference with its
`make debug` seed, see oracle/Makefile) on seeded synthetic files and records the SHA-256 of the `.qvz` it writes and
of its `-u` dump, plus the rate / distortion / size fields of its `-s` line.  The inputs come from
qvz_b200.synth.synth_rows on the CPU generator, so tests/test_cli_gpu.py regenerates the very same bytes on the GPU
box (where neither /root/reference nor minutes of single-threaded reference time are available) and compares what
the new command line and the reference-side binding produce with these hashes.

    python tests/golden/make_cli_golden.py [case ...]        # writes tests/golden/cli_reference_hashes.json
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qvz_b200.synth import synth_rows  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "qvz_ref_det")
OUT = os.path.join(ROOT, "tests", "golden", "cli_reference_hashes.json")

# name -> (lines, columns, synth profile, synth seed, flags)      [BASELINE.json configs at their read lengths]
CASES = {
    "cfg2_shape": (30_000, 150, "illumina", 7102, ["-r", "2", "-d", "L", "-c", "1"]),
    "cfg3_shape": (30_000, 150, "illumina", 7103, ["-f", "0.5", "-d", "A", "-c", "3", "-T", "4"]),
    "cfg4_shape": (24_000, 150, "illumina", 7104, ["-f", "1.0", "-d", "M", "-c", "5"]),
    "cfg5_shape": (20_000, 250, "miseq", 7105, ["-r", "4", "-d", "M", "-c", "2"]),
    "cfg2_full": (20_000_000, 150, "illumina", 1234, ["-r", "2", "-d", "L", "-c", "1"]),
}


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    names = sys.argv[1:] or [n for n in CASES if n != "cfg2_full"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        n, c, profile, seed, flags = CASES[name]
        with tempfile.TemporaryDirectory(dir=os.environ.get("QVZ_TMP")) as d:
            src, dst, uf = os.path.join(d, "in.txt"), os.path.join(d, "out.qvz"), os.path.join(d, "u.txt")
            t0 = time.time()
            synth_rows(n, c, seed=seed, profile=profile).numpy().tofile(src)
            t1 = time.time()
            r = subprocess.run([REF] + flags + ["-u", uf, "-s", src, dst], capture_output=True, text=True, check=True)
            t2 = time.time()
            f = [x.strip() for x in r.stdout.strip().split(",")]
            res[name] = {"lines": n, "columns": c, "profile": profile, "seed": seed, "flags": flags,
                         "input_sha256": sha(src), "qvz_sha256": sha(dst), "qvz_bytes": os.path.getsize(dst),
                         "u_sha256": sha(uf), "rate": f[1], "distortion": f[3], "size": f[7],
                         "reference_seconds": round(t2 - t1, 1), "synth_seconds": round(t1 - t0, 1)}
            print(name, res[name], flush=True)
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
