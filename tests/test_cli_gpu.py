"""The drop-in command line (qvz_b200/host/qvz) on a B200 against the reference's own binary.

north_star's correctness definition: byte-identical `-u` dumps and `.qvz` files versus the reference on identical
input.  oracle/_ref/qvz_ref_det is the unmodified reference built with its `make debug` seed semantics
(-DDEBUG: WELL seed 0x55555555, src/qv_stream.c:79-83); the new CLI selects the same seed with QVZ_DEBUG_SEED=1.
Both processes consume libc rand() in the same order for the initial centroids (src/cluster.c:199-200)."""
import os
import subprocess

import numpy as np
import pytest

from qvz_b200.synth import synth_rows
from tests.conftest import GOLDEN_CASES, load_golden

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "qvz_b200", "host", "qvz")
REF = os.path.join(ROOT, "oracle", "_ref", "qvz_ref_det")
DIST = {1: "A", 2: "M", 3: "L"}


def _run(exe, args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([exe] + args, capture_output=True, text=True, env=env, timeout=600)


def _mode_args(mode, ratio):
    return ["-f" if int(mode) == 0 else "-r", repr(float(ratio))]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cli_reproduces_golden_file(tmp_path, name):
    """The committed fixtures hold the bytes the reference's encode() wrote for these inputs."""
    g = load_golden(name)
    src, dst, uf = str(tmp_path / "in.txt"), str(tmp_path / "out.qvz"), str(tmp_path / "u.txt")
    g["rows"].tofile(src)
    args = _mode_args(g["mode"], g["ratio"]) + ["-d", DIST[int(g["dist"])], "-c", str(g["clusters"]), "-T", str(int(g["threshold"])),
                                                "-u", uf, "-s", src, dst]
    r = _run(CLI, args, {"QVZ_DEBUG_SEED": "1"})
    assert r.returncode == 0, r.stdout + r.stderr
    assert np.array_equal(np.fromfile(dst, np.uint8), g["qvz"])
    assert np.array_equal(np.fromfile(uf, np.uint8).reshape(g["qv"].shape), g["qv"])
    assert r.stdout.startswith("rate, ")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/qvz_ref_det not built")
@pytest.mark.parametrize("n,c,flags,gpus", [(60_000, 150, ["-f", "0.5", "-d", "A", "-c", "3", "-T", "4"], 1),   # cfg3's command line and read length
                                            (30_000, 36, ["-f", "0.5", "-d", "A", "-c", "3", "-T", "4"], 1),
                                            (20_000, 50, ["-r", "2", "-d", "L", "-c", "1"], 1),
                                            (25_000, 30, ["-f", "1.0", "-d", "M", "-c", "2"], 2)])
def test_cli_vs_reference_binary(tmp_path, n, c, flags, gpus):
    import torch
    if gpus > torch.cuda.device_count():
        gpus = 1                                   # QVZ_GPUS shards over devices 0..n-1; one device still exercises the path below
    rows = synth_rows(n, c, seed=4242 + n).numpy()
    src = str(tmp_path / "in.txt")
    rows.tofile(src)
    out = {}
    for tag, exe, env in (("ref", REF, {}), ("new", CLI, {"QVZ_DEBUG_SEED": "1", "QVZ_GPUS": str(gpus)})):
        dst, uf = str(tmp_path / f"{tag}.qvz"), str(tmp_path / f"{tag}.u")
        r = _run(exe, flags + ["-u", uf, "-s", src, dst], env)
        assert r.returncode == 0, r.stdout + r.stderr
        out[tag] = (np.fromfile(dst, np.uint8), np.fromfile(uf, np.uint8), r.stdout)
    assert np.array_equal(out["new"][1], out["ref"][1]), "-u dump differs"
    assert out["new"][0].size == out["ref"][0].size and np.array_equal(out["new"][0], out["ref"][0]), ".qvz differs"
    # the -s line: same rate, distortion and size fields (the time field differs, of course)
    f_new, f_ref = out["new"][2].split(","), out["ref"][2].split(",")
    assert [f_new[i].strip() for i in (1, 3, 7)] == [f_ref[i].strip() for i in (1, 3, 7)]
    # the reference's own round-trip criterion (test.sh:7-9): its decoder reproduces the -u dump from OUR file
    dec = str(tmp_path / "dec.txt")
    r = _run(REF, ["-x", str(tmp_path / "new.qvz"), dec])
    assert r.returncode == 0
    assert np.array_equal(np.fromfile(dec, np.uint8), out["new"][1])
    # ... and the new command line's -x decodes the REFERENCE's file to the same lines
    dec2 = str(tmp_path / "dec2.txt")
    r = _run(CLI, ["-x", str(tmp_path / "ref.qvz"), dec2])
    assert r.returncode == 0, r.stdout + r.stderr
    assert np.array_equal(np.fromfile(dec2, np.uint8), out["ref"][1])


def test_cli_errors(tmp_path):
    r = _run(CLI, [])
    assert r.returncode == 1 and "Missing required filenames." in r.stdout
    r = _run(CLI, ["-f", "0.5", str(tmp_path / "missing.txt"), str(tmp_path / "o.qvz")])
    assert r.returncode == 1 and "load_file returned error" in r.stdout
    r = _run(CLI, ["-h"])
    assert r.returncode == 0 and r.stdout.startswith("Usage:")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/qvz_ref_det not built")
def test_cfg1_full_size_vs_reference_binary(tmp_path):
    """BASELINE.json configs[0] at its full size: `qvz -q -f 1.0 -d M -c 1` on 1M synthetic 100-bp lines -- the one
    config the reference itself finishes in about a minute.  Same .qvz bytes, same -u dump, same -s numbers."""
    n, c = 1_000_000, 100
    rows = synth_rows(n, c, seed=1234, device="cuda").cpu().numpy()
    src = str(tmp_path / "cfg1.txt")
    rows.tofile(src)
    flags = ["-q", "-f", "1.0", "-d", "M", "-c", "1"]
    out = {}
    for tag, exe, env in (("ref", REF, {}), ("new", CLI, {"QVZ_DEBUG_SEED": "1"})):
        dst, uf = str(tmp_path / f"{tag}.qvz"), str(tmp_path / f"{tag}.u")
        r = _run(exe, flags + ["-u", uf, "-s", src, dst], env)
        assert r.returncode == 0, r.stdout + r.stderr
        out[tag] = (np.fromfile(dst, np.uint8), np.fromfile(uf, np.uint8), r.stdout)
    assert np.array_equal(out["new"][1], out["ref"][1]), "-u dump differs"
    assert out["new"][0].size == out["ref"][0].size and np.array_equal(out["new"][0], out["ref"][0]), ".qvz differs"
    f_new, f_ref = out["new"][2].split(","), out["ref"][2].split(",")
    assert [f_new[i].strip() for i in (1, 3, 7)] == [f_ref[i].strip() for i in (1, 3, 7)]
    assert float(f_new[5]) < float(f_ref[5])          # wall time of the whole command, reference vs new
