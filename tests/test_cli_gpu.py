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
# the reference's own main.c / codebook design / coder linked against the GPU front end (integration/Makefile)
REF_GPU = os.path.join(ROOT, "integration", "_build", "qvz_ref_gpu")
HASHES = os.path.join(ROOT, "tests", "golden", "cli_reference_hashes.json")
DIST = {1: "A", 2: "M", 3: "L"}


def _sha(path):
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def _reference_hashes():
    import json
    return json.load(open(HASHES)) if os.path.exists(HASHES) else {}


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


@pytest.mark.parametrize("name", ["cfg2_shape", "cfg3_shape", "cfg4_shape", "cfg5_shape", "cfg2_full"])
def test_cli_vs_cached_reference_hashes(tmp_path, name):
    """BASELINE.json's command lines at their real read lengths (150 / 250 columns, up to 5 clusters), and configs[1] at
    its FULL size (20 M x 150): the `.qvz` and the `-u` dump must hash to what the unmodified reference binary produced
    for the same seeded input (tests/golden/make_cli_golden.py ran it once; the reference needs minutes to hours of one
    CPU core for these -- test.sh:7-9 is its own criterion).  Both front ends are checked: the new command line, and the
    reference's own main.c linked against libqvz_gpu.so (integration/gpu_frontend.c)."""
    want = _reference_hashes().get(name)
    if not want:
        pytest.skip(f"no cached reference hashes for {name}")
    big = want["lines"] > 1_000_000
    if big and os.environ.get("QVZ_SKIP_FULL"):
        pytest.skip("QVZ_SKIP_FULL is set (the full-size case regenerates 3 GB of input on the CPU generator)")
    src = str(tmp_path / "in.txt")
    synth_rows(want["lines"], want["columns"], seed=want["seed"], profile=want["profile"]).numpy().tofile(src)
    assert _sha(src) == want["input_sha256"], "the generator no longer reproduces the cached input"
    exes = [("new", CLI, {"QVZ_DEBUG_SEED": "1"})]
    if os.path.exists(REF_GPU) and name == "cfg3_shape":           # (its codebook design is the reference's own: minutes for the others)
        exes.append(("ref_gpu", REF_GPU, {}))
    for tag, exe, env in exes:
        dst, uf = str(tmp_path / f"{tag}.qvz"), str(tmp_path / f"{tag}.u")
        r = _run(exe, want["flags"] + ["-u", uf, "-s", src, dst], env)
        assert r.returncode == 0, r.stdout + r.stderr
        assert _sha(uf) == want["u_sha256"], f"{tag}: -u dump differs from the reference's"
        assert os.path.getsize(dst) == want["qvz_bytes"] and _sha(dst) == want["qvz_sha256"], f"{tag}: .qvz differs from the reference's"
        f = [x.strip() for x in r.stdout.strip().split(",")]
        assert (f[1], f[3], f[7]) == (want["rate"], want["distortion"], want["size"]), tag
        os.remove(dst)
        os.remove(uf)


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(REF_GPU)), reason="reference binaries not built")
@pytest.mark.parametrize("n,c,flags", [(40_000, 100, ["-f", "0.5", "-d", "M", "-c", "3", "-T", "4", "-v"]),
                                       (15_000, 61, ["-r", "3", "-d", "L", "-c", "2"]),
                                       (9_000, 37, ["-f", "0.8", "-d", "A", "-c", "1", "-v"])])
def test_reference_main_on_gpu_front_end(tmp_path, n, c, flags):
    """The real drop-in: the reference's unmodified main.c, load_file, generate_codebooks, write_codebooks and arithmetic
    coder, with do_kmeans_clustering / calculate_statistics / start_qv_compression provided by integration/gpu_frontend.c
    on top of the C ABI.  Same `.qvz`, same `-u` dump, same stdout (minus timings) as the all-CPU reference binary."""
    rows = synth_rows(n, c, seed=99 + n).numpy()
    src = str(tmp_path / "in.txt")
    rows.tofile(src)
    out = {}
    for tag, exe in (("ref", REF), ("gpu", REF_GPU)):
        dst, uf = str(tmp_path / f"{tag}.qvz"), str(tmp_path / f"{tag}.u")
        r = _run(exe, flags + ["-u", uf, "-s", src, dst])
        assert r.returncode == 0, r.stdout + r.stderr
        text = [ln for ln in r.stdout.replace(str(tmp_path / f"{tag}."), "X.").splitlines() if " seconds" not in ln and not ln.startswith("rate, ")]      # timings differ, of course
        stats = [x.strip() for x in r.stdout.strip().splitlines()[-1].split(",")]
        out[tag] = (np.fromfile(dst, np.uint8), np.fromfile(uf, np.uint8), text, (stats[1], stats[3], stats[7]))
    assert np.array_equal(out["gpu"][1], out["ref"][1]), "-u dump differs"
    assert out["gpu"][0].size == out["ref"][0].size and np.array_equal(out["gpu"][0], out["ref"][0]), ".qvz differs"
    assert out["gpu"][2] == out["ref"][2], "stdout differs"
    assert out["gpu"][3] == out["ref"][3]
