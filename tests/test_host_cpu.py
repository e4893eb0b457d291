"""Host side (qvz_b200/host, include/qvz_host.h) against the reference, no GPU needed.

The golden fixtures hold what the UNMODIFIED reference produced for three small files: the conditional counts,
the flattened cond_quantizer_list_t of generate_codebooks, the symbol stream of the quantize walk and the bytes
of the .qvz file written by its encode().  Feeding the reference's counts to qvz_host_design must reproduce its
tables exactly (same doubles, same Lloyd-Max decisions), and feeding its symbol stream to qvz_host_encode must
reproduce the .qvz file byte for byte (codebook text, seed, arithmetic-coded stream)."""
import os
import re

import numpy as np
import pytest

from oracle.bindings import DEBUG_SEED, DIST_LORENTZ, DIST_MANHATTAN, DIST_MSE, MODE_FIXED, MODE_RATIO, kmeans_init_lines
from qvz_b200 import hostlib
from qvz_b200.synth import synth_rows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    hostlib.build()


def test_exports_match_header():
    import ctypes
    header = open(os.path.join(ROOT, "include", "qvz_host.h")).read()
    declared = set(re.findall(r"\b(qvz_host_[a-z0-9_]+)\s*\(", header))
    assert declared == set(hostlib.EXPORTS)
    L = ctypes.CDLL(hostlib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_distortion_matrices(golden):
    d = hostlib.distortion_matrix(int(golden["dist"]))
    assert np.array_equal(d, golden["t_distortion"])                        # bit-exact doubles (log2 included)


def _design(g):
    return hostlib.design_codebooks(g["cond_counts"], g["columns"], g["clusters"], int(g["mode"]), float(g["ratio"]),
                                    int(g["dist"]))


def test_design_reproduces_reference_tables(golden):
    cb = _design(golden)
    for name in ("nctx", "ctx_of", "q_off", "qratio", "qmap", "smap"):
        assert np.array_equal(getattr(cb, name), golden["t_" + name]), name


@pytest.fixture(params=["sequential", "pipelined"])
def coder_mode(request, monkeypatch):
    """qvz_host_encode codes small inputs on one thread and large ones through its model/interval pipeline;
    the tests force either path (and short blocks, so that the small fixtures span many of them)."""
    if request.param == "pipelined":
        monkeypatch.setenv("QVZ_CODER_THREADS", "4")
        monkeypatch.setenv("QVZ_CODER_BLOCK", "97")
    else:
        monkeypatch.setenv("QVZ_CODER_THREADS", "1")
    return request.param


def test_container_bytes_equal_reference_file(golden, tmp_path, coder_mode):
    g = golden
    cb = _design(g)
    n = g["rows"].shape[0]
    head = cb.codebook_bytes(n)
    qvz = g["qvz"]
    assert np.array_equal(head, qvz[:head.size])                             # 9-byte header + codebook text
    path = str(tmp_path / "out.qvz")
    written = cb.encode(path, g["ids"], g["symbols"], DEBUG_SEED)
    mine = np.fromfile(path, np.uint8)
    assert mine.size == qvz.size and np.array_equal(mine, qvz)               # the whole file, byte for byte
    assert written == qvz.size - head.size - 128                             # what start_qv_compression returns


def test_encode_rejects_malformed_symbols(golden, tmp_path, coder_mode):
    g = golden
    cb = _design(g)
    bad = g["symbols"].copy()
    bad[3, 2] = 0x7F                                                         # a state no quantizer has
    with pytest.raises(RuntimeError):
        cb.encode(str(tmp_path / "bad.qvz"), g["ids"], bad, DEBUG_SEED)


@pytest.mark.parametrize("n,c,k,mode,ratio,dist", [(3000, 30, 2, MODE_RATIO, 1.0, DIST_MSE), (2000, 25, 1, MODE_FIXED, 3.0, DIST_LORENTZ),
                                                   (2500, 18, 3, MODE_RATIO, 0.0, DIST_MANHATTAN), (1500, 40, 1, MODE_FIXED, 0.5, DIST_MSE)])
def test_design_and_file_vs_compiled_reference(ref, tmp_path, n, c, k, mode, ratio, dist, monkeypatch):
    # fresh inputs through the unmodified reference (oracle/_ref): its counts -> my design == its tables;
    # its encode() of the same file == my container from its symbol stream
    rows = synth_rows(n, c, seed=900 + n).numpy()
    picks = kmeans_init_lines(n, k, ref.rand_stream(2 * k))
    s = ref.session(rows, c, k, mode=mode, ratio=ratio, distortion=dist)
    ids = s.kmeans(picks)["ids"]
    counts, _ = s.stats()
    t = s.tables()
    cb = hostlib.design_codebooks(counts, c, k, mode, ratio, dist)
    for name in ("nctx", "ctx_of", "q_off", "qratio", "qmap", "smap"):
        assert np.array_equal(getattr(cb, name), getattr(t, name)), name
    q = s.quantize(DEBUG_SEED)
    src, dst = str(tmp_path / "in.txt"), str(tmp_path / "ref.qvz")
    rows.tofile(src)
    ref.encode_file(src, dst, None, clusters=k, mode=mode, ratio=ratio, distortion=dist)
    mine = str(tmp_path / "mine.qvz")
    for threads, block in (("1", "8192"), ("4", "97"), ("2", "1000")):           # one thread, and the model/interval pipeline
        monkeypatch.setenv("QVZ_CODER_THREADS", threads)
        monkeypatch.setenv("QVZ_CODER_BLOCK", block)
        cb.encode(mine, ids, q["symbols"], DEBUG_SEED)
        assert np.array_equal(np.fromfile(mine, np.uint8), np.fromfile(dst, np.uint8)), threads


def test_decode_reproduces_u_dump(golden, tmp_path):
    """qvz -x on the reference's own .qvz file gives the reference's -u dump (the reference's test.sh criterion)."""
    src, dst = str(tmp_path / "in.qvz"), str(tmp_path / "out.txt")
    golden["qvz"].tofile(src)
    assert hostlib.decode_file(src, dst) == golden["rows"].shape[0]
    assert np.array_equal(np.fromfile(dst, np.uint8).reshape(golden["qv"].shape), golden["qv"])


def test_decode_rejects_garbage(tmp_path):
    src = str(tmp_path / "bad.qvz")
    np.arange(200, dtype=np.uint8).tofile(src)
    with pytest.raises(RuntimeError):
        hostlib.decode_file(src, str(tmp_path / "o.txt"))


def test_custom_distortion_file(tmp_path):
    """-D FILE (gen_custom_distortion, src/distortion.c:100-145): row x of the file is D[x + 72*y] for y = 0..71."""
    import ctypes as C
    m = np.round(np.random.default_rng(9).random((72, 72)) * 5, 3)       # short fields: the reference reads rows through a 1024-byte buffer
    path = tmp_path / "dist.csv"
    with open(path, "w") as f:
        f.write("# comment line\n")
        for x in range(72):
            f.write(",".join(repr(float(v)) for v in m[x]) + "\n")
    out = np.zeros(72 * 72, np.float64)
    L = hostlib.load()
    assert L.qvz_host_distortion_file(str(path).encode(), out.ctypes.data_as(C.POINTER(C.c_double))) == 0
    assert np.array_equal(out.reshape(72, 72), m.T)                       # index x + 72*y
    assert L.qvz_host_distortion_file(b"/nonexistent/file", out.ctypes.data_as(C.POINTER(C.c_double))) == -1
