"""CUDA path (through the C ABI of include/qvz_gpu.h) against the oracle, the golden fixtures and the
compiled reference.  Bit-exact everywhere: integer/byte/index work, and the per-line distortion doubles
(same additions in the same order)."""
import numpy as np
import pytest

from oracle.bindings import DEBUG_SEED, kmeans_init_lines, ref_available
from qvz_b200.synth import synth_rows
from tests.helpers import synthetic_tables

pytestmark = pytest.mark.gpu

GLIBC_RAND = [1804289383, 846930886, 1681692777, 1714636915, 1957747793, 424238335, 719885386, 1649760492,
              596516649, 1189641421, 1025202362, 1350490027, 783368690, 1102520059, 2044897763, 1967513926]


@pytest.fixture(scope="module")
def gpu():
    from qvz_b200 import lib
    h = lib.Handle(0)
    yield h
    h.close()


def _load(gpu, rows, c):
    gpu.load_rows(rows, rows.shape[0], c, rows.shape[1])


def _quantize_both_paths(gpu, tables, seed):
    """The batched shared-memory walk (default) and the line-major walk must agree bit for bit."""
    import os
    q = gpu.quantize(tables, seed, want_qv=True, want_err=True)
    os.environ["QVZ_FORCE_LINE_MAJOR"] = "1"
    try:
        q2 = gpu.quantize(tables, seed, want_qv=True, want_err=True)
    finally:
        del os.environ["QVZ_FORCE_LINE_MAJOR"]
    for key in ("symbols", "qv", "line_err"):
        assert np.array_equal(q[key], q2[key]), "paths differ: " + key
    return q


def test_golden_all_stages(gpu, golden):
    g = golden
    c, K = g["columns"], g["clusters"]
    _load(gpu, g["rows"], c)
    init = g["rows"][g["picks"].astype(np.int64), :c]
    r = gpu.kmeans(init, float(g["threshold"]))
    assert r["iters"] == int(g["iters"])
    assert np.array_equal(r["ids"], g["ids"])
    assert np.array_equal(r["means"], g["means"])
    assert np.array_equal(r["counts"], g["kcounts"])
    assert np.array_equal(r["moved"], g["moved"])
    assert np.array_equal(gpu.cond_counts(), g["cond_counts"])
    q = _quantize_both_paths(gpu, g["tables"], DEBUG_SEED)
    assert np.array_equal(q["symbols"], g["symbols"])
    assert np.array_equal(q["qv"], g["qv"])                       # the -u image, byte for byte
    assert np.array_equal(q["line_err"], g["line_err"])           # bit-exact doubles
    assert q["line_err"].sum() / g["rows"].shape[0] == pytest.approx(float(g["distortion"]), rel=1e-12)


@pytest.mark.parametrize("words", [0, 1, 31, 32, 33, 1000, 1_000_000, 123_456_789])
def test_well_jump(gpu, oracle, words):
    seed = np.random.default_rng(words).integers(0, 2**32, 32, dtype=np.uint32)
    assert np.array_equal(gpu.well_jump(seed, words), oracle.well_state_after(seed, words))


def test_well_jump_known_answer(gpu):
    assert int(gpu.well_jump(DEBUG_SEED, 1_000_000)[0]) == 0x90e10060
    # 64-bit offsets: A^(a+b) = A^a A^b
    a, b = 7_500_000_000, 29_999_999_999
    assert np.array_equal(gpu.well_jump(gpu.well_jump(DEBUG_SEED, a), b), gpu.well_jump(DEBUG_SEED, a + b))


@pytest.mark.parametrize("n,c,k,thr", [(200_000, 150, 3, 4.0), (50_000, 100, 1, 4.0), (30_001, 37, 5, 0.0),
                                       (4099, 250, 2, 4.0), (999, 5, 8, 4.0), (2000, 9, 11, 4.0),
                                       (1, 7, 1, 4.0), (5, 1, 1, 4.0), (300, 1022, 2, 4.0)])
def test_kmeans_and_counts_vs_oracle(gpu, oracle, n, c, k, thr):
    rows = synth_rows(n, c, seed=1000 + n + c).numpy()
    picks = kmeans_init_lines(n, k, GLIBC_RAND) if 2 * k <= len(GLIBC_RAND) else [(i * 7919 + 13) % n for i in range(k)]
    if len(set(picks)) < k:                       # tiny inputs: make the initial rows distinct
        picks = list(range(k))
    init = rows[picks, :c]
    o = oracle.kmeans(rows, c, init, thr)
    _load(gpu, rows, c)
    if o["iters"] < 0:
        from qvz_b200.lib import QvzError
        with pytest.raises(QvzError):
            gpu.kmeans(init, thr)
        return
    r = gpu.kmeans(init, thr)
    assert r["iters"] == o["iters"]
    for key in ("ids", "means", "counts", "moved"):
        assert np.array_equal(r[key], o[key]), key
    assert np.array_equal(gpu.cond_counts(), oracle.cond_counts(rows, c, k, o["ids"]))


@pytest.mark.parametrize("shape", ["2128", "264", "2256", "464", "1256", "sorted"])
@pytest.mark.parametrize("n,c,k", [(70_000, 150, 5), (40_000, 250, 2), (33_333, 101, 8)])
def test_kmeans_kernel_variants(gpu, oracle, n, c, k, shape):
    """Every tile shape of the register-blocked k-means kernel (rows per thread x threads: tensor-core column sums,
    full and incremental iterations) and the counting-sort kernel give the oracle's ids, means, `moved` log and counts."""
    import os
    rows = synth_rows(n, c, seed=77 + n).numpy()
    picks = kmeans_init_lines(n, k, GLIBC_RAND)
    init = rows[picks, :c]
    o = oracle.kmeans(rows, c, init, 4.0)
    assert o["iters"] >= 2                         # an incremental iteration has run
    _load(gpu, rows, c)
    key, val = ("QVZ_KM_SORTED", "1") if shape == "sorted" else ("QVZ_KM_SHAPE", shape)
    os.environ[key] = val
    try:
        r = gpu.kmeans(init, 4.0)
    finally:
        del os.environ[key]
    assert r["iters"] == o["iters"]
    for name in ("ids", "means", "counts", "moved"):
        assert np.array_equal(r[name], o[name]), name


@pytest.mark.parametrize("n,c,k,dist", [(50_000, 150, 5, "M"), (12_345, 250, 2, "A"), (8000, 101, 1, "L"), (5, 3, 2, "M")])
def test_quantize_without_draws(gpu, oracle, n, c, k, dist):
    """Tables in which no context mixes its two quantizers (qratio 0 or 128 only): the walk neither generates nor reads
    the draw stream, and gives the oracle's bytes all the same; so does the same walk forced to read the draws."""
    import os
    rows = synth_rows(n, c, seed=3000 + n).numpy()
    ids = np.random.default_rng(n).integers(0, k, n, dtype=np.uint8)
    t = synthetic_tables(k, c, seed=n, dist=dist, mixing=False)
    seed = np.random.default_rng(c).integers(0, 2**31, 32, dtype=np.uint32)
    _load(gpu, rows, c)
    gpu.set_clusters(k, ids)
    o = oracle.quantize(rows, c, ids, t, seed)
    q = _quantize_both_paths(gpu, t, seed)
    assert gpu.timings()["quantize_draws_ms"] >= 0.0
    os.environ["QVZ_FORCE_DRAWS"] = "1"
    try:
        q2 = gpu.quantize(t, seed, want_qv=True, want_err=True)
    finally:
        del os.environ["QVZ_FORCE_DRAWS"]
    for res in (q, q2):
        assert np.array_equal(res["symbols"], o["symbols"])
        assert np.array_equal(res["qv"], o["qv"])
        assert np.array_equal(res["line_err"], o["line_err"])


@pytest.mark.parametrize("n,c,k,dist", [(60_000, 150, 3, "L"), (20_000, 101, 1, "M"), (7777, 250, 5, "A"), (6, 3, 2, "L")])
def test_quantize_vs_oracle_synthetic_tables(gpu, oracle, n, c, k, dist):
    rows = synth_rows(n, c, seed=2000 + n).numpy()
    ids = np.random.default_rng(n).integers(0, k, n, dtype=np.uint8)
    t = synthetic_tables(k, c, seed=n, dist=dist)
    seed = np.random.default_rng(c).integers(0, 2**31, 32, dtype=np.uint32)
    _load(gpu, rows, c)
    gpu.set_clusters(k, ids)
    q = _quantize_both_paths(gpu, t, seed)
    o = oracle.quantize(rows, c, ids, t, seed)
    assert np.array_equal(q["symbols"], o["symbols"])
    assert np.array_equal(q["qv"], o["qv"])
    assert np.array_equal(q["line_err"], o["line_err"])


def test_quantize_shard_offset(gpu, oracle):
    # a shard whose first line is L0 continues the draw stream at draw L0*C (multi-GPU contract)
    n, c, k, L0 = 30_000, 150, 2, 12_344
    rows = synth_rows(n, c, seed=5).numpy()
    ids = np.random.default_rng(1).integers(0, k, n, dtype=np.uint8)
    t = synthetic_tables(k, c, seed=9)
    whole = oracle.quantize(rows, c, ids, t, DEBUG_SEED)
    tail = np.ascontiguousarray(rows[L0:])
    gpu.load_rows(tail, n - L0, c, c + 1, first_line=L0)
    gpu.set_clusters(k, ids[L0:])
    q = _quantize_both_paths(gpu, t, DEBUG_SEED)
    assert np.array_equal(q["symbols"], whole["symbols"][L0:])
    assert np.array_equal(q["qv"], whole["qv"][L0:])
    assert np.array_equal(q["line_err"], whole["line_err"][L0:])


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_against_compiled_reference(gpu, ref):
    # the unmodified reference functions (do_kmeans_clustering loop, calculate_statistics, generate_codebooks,
    # choose_quantizer walk) on the same seeded input, C small so codebook design takes seconds
    n, c, k = 40_000, 24, 3
    rows = synth_rows(n, c, seed=31).numpy()
    picks = kmeans_init_lines(n, k, ref.rand_stream(2 * k))
    s = ref.session(rows, c, k, threshold=4.0, ratio=0.7)
    km = s.kmeans(picks)
    _load(gpu, rows, c)
    r = gpu.kmeans(rows[picks, :c], 4.0)
    assert r["iters"] == km["iters"] and np.array_equal(r["ids"], km["ids"]) and np.array_equal(r["means"], km["means"])
    counts, _ = s.stats()
    assert np.array_equal(gpu.cond_counts(), counts)
    t = s.tables()
    rq = s.quantize(DEBUG_SEED)
    q = _quantize_both_paths(gpu, t, DEBUG_SEED)
    assert np.array_equal(q["symbols"], rq["symbols"])
    assert np.array_equal(q["qv"], rq["qv"])
    assert np.array_equal(q["line_err"], rq["line_err"])


def test_error_paths(gpu):
    from qvz_b200.lib import QvzError
    rows = synth_rows(64, 8, seed=1).numpy()
    bad = rows.copy()
    bad[5, 3] = 33 + 72                           # out of the 72-symbol alphabet
    with pytest.raises(QvzError) as e:
        gpu.load_rows(bad, 64, 8, 9)
    assert e.value.code == 4
    _load(gpu, rows, 8)
    with pytest.raises(QvzError) as e:            # identical initial centroids -> empty cluster (reference: SIGFPE)
        gpu.kmeans(np.stack([rows[0, :8], rows[0, :8]]), 4.0)
    assert e.value.code == 3
    gpu.set_clusters(1, np.zeros(64, np.uint8))
    t = synthetic_tables(1, 8, seed=3)
    t.ctx_of[8 * 72 * 0 + 72 * 3: 72 * 4] = 0xFF   # column 3 loses all its contexts (reference: assert)
    with pytest.raises(QvzError) as e:
        gpu.quantize(t, DEBUG_SEED)
    assert e.value.code == 5
    with pytest.raises(QvzError):                 # shards must start on a WELL word boundary
        gpu.load_rows(rows, 64, 8, 9, first_line=2)


def test_sharded_equals_whole(gpu, oracle):
    """Two shards (two handles on this GPU) driven through the stepping interface, with the all-reduce of
    the integer sums emulated by adding the two device buffers: every result equals the whole-file run."""
    import torch
    from qvz_b200 import lib
    from qvz_b200.dist import shard_bounds
    n, c, k, thr = 50_003, 37, 3, 4.0
    rows = synth_rows(n, c, seed=41).numpy()
    picks = [(i * 7919 + 13) % n for i in range(k)]
    init = rows[picks, :c]
    whole = oracle.kmeans(rows, c, init, thr)
    b = shard_bounds(n, 2)
    hs = [lib.Handle(0), lib.Handle(0)]
    try:
        for r, h in enumerate(hs):
            h.load_rows(np.ascontiguousarray(rows[b[r]:b[r + 1]]), b[r + 1] - b[r], c, c + 1, first_line=b[r])
            h.kmeans_begin(init)
        sums = [torch.zeros(k * c + k, dtype=torch.int64, device="cuda") for _ in hs]
        iters, loop = 0, True
        while loop and iters < 1000:
            for h, s in zip(hs, sums):
                h.kmeans_assign_dev(s.data_ptr())
            torch.cuda.synchronize()
            tot = sums[0] + sums[1]
            for s in sums:
                s.copy_(tot)
            torch.cuda.synchronize()
            moved = [h.kmeans_update_dev(s.data_ptr())[0] for h, s in zip(hs, sums)]
            assert np.array_equal(moved[0], moved[1])
            loop = moved[0].max() > thr
            iters += 1
        assert iters == whole["iters"]
        ids = np.concatenate([h.kmeans_end()[0] for h in hs])
        assert np.array_equal(ids, whole["ids"])
        cnt = [torch.zeros(h.cond_counts_len(), dtype=torch.int32, device="cuda") for h in hs]
        for h, t in zip(hs, cnt):
            h.cond_counts_dev(t.data_ptr())
        torch.cuda.synchronize()
        got = (cnt[0] + cnt[1]).cpu().numpy().view(np.uint32).reshape(k, 1 + 72 * (c - 1), 72)
        assert np.array_equal(got, oracle.cond_counts(rows, c, k, whole["ids"]))
        t = synthetic_tables(k, c, seed=2)
        wq = oracle.quantize(rows, c, whole["ids"], t, DEBUG_SEED)
        q = [h.quantize(t, DEBUG_SEED, want_qv=True, want_err=True) for h in hs]
        for key in ("symbols", "qv", "line_err"):
            assert np.array_equal(np.concatenate([x[key] for x in q]), wq[key]), key
    finally:
        for h in hs:
            h.close()


def test_prefetched_draws_and_host_reduction_entry_points(gpu, oracle):
    """qvz_gpu_prefetch_draws (draws generated early on the auxiliary stream) and the host-sum stepping calls
    (qvz_gpu_kmeans_assign_host / _update_host) give the same bits as the plain calls."""
    import ctypes as C
    from qvz_b200.lib import f64p, i64p, u32p
    n, c, k = 40_003, 61, 3
    rows = synth_rows(n, c, seed=77).numpy()
    init = rows[[5, 20_000, 39_999], :c]
    o = oracle.kmeans(rows, c, init, 4.0)
    _load(gpu, rows, c)
    gpu.prefetch_draws(DEBUG_SEED)                     # before the clusters or the tables exist
    gpu.kmeans_begin(init)
    sums = np.zeros(k * c + k, np.int64)
    moved, counts = np.zeros(k), np.zeros(k, np.uint32)
    iters, loop = 0, True
    while loop:
        assert gpu.L.qvz_gpu_kmeans_assign_host(gpu.h, sums.ctypes.data_as(i64p)) == 0
        assert gpu.L.qvz_gpu_kmeans_update_host(gpu.h, sums.ctypes.data_as(i64p), moved.ctypes.data_as(f64p), counts.ctypes.data_as(u32p)) == 0
        assert np.array_equal(moved, o["moved"][iters])
        loop = moved.max() > 4.0
        iters += 1
    assert iters == o["iters"] and np.array_equal(counts, o["counts"])
    ids, means = gpu.kmeans_end()
    assert np.array_equal(ids, o["ids"]) and np.array_equal(means, o["means"])
    t = synthetic_tables(k, c, seed=8)
    q = gpu.quantize(t, DEBUG_SEED, want_qv=True, want_err=True)          # consumes the prefetched draws
    q2 = gpu.quantize(t, DEBUG_SEED, want_qv=True, want_err=True)         # generates them again
    r = oracle.quantize(rows, c, o["ids"], t, DEBUG_SEED)
    for key in ("symbols", "qv", "line_err"):
        assert np.array_equal(q[key], r[key]) and np.array_equal(q2[key], r[key]), key
    other = np.random.default_rng(3).integers(0, 2**31, 32, dtype=np.uint32)
    gpu.prefetch_draws(other)                          # a prefetch for a different seed must not be used
    q3 = gpu.quantize(t, DEBUG_SEED)
    assert np.array_equal(q3["symbols"], r["symbols"])


def test_cfg_scale_properties(gpu):
    """A cfg3-shaped run far beyond what the CPU oracle checks in seconds (2M x 150, K = 3, -f 0.5 -d A), verified through
    size-independent properties: every line is counted exactly once per column, the count tables agree with the
    cluster sizes, re-running is idempotent, the symbols are valid states of the quantizer the context selects and
    the host coder accepts the whole stream."""
    import tempfile
    import torch
    from qvz_b200 import hostlib
    n, c, k = 2_000_000, 150, 3
    rows = synth_rows(n, c, seed=2024, device="cuda").cpu().numpy()
    init = rows[[930_886, 636_915, 238_335], :c]
    _load(gpu, rows, c)
    km = gpu.kmeans(init, 4.0)
    assert km["counts"].sum() == n and np.array_equal(np.bincount(km["ids"], minlength=k), km["counts"])
    counts = gpu.cond_counts()
    per_col = counts.reshape(k, -1).sum(1)
    assert np.array_equal(per_col, km["counts"].astype(np.uint64) * c)           # each line once per column, in its cluster
    assert np.array_equal(counts[:, 0, :].sum(1), km["counts"])                  # column 0 rows
    assert np.array_equal(gpu.cond_counts(), counts)                             # idempotent
    km2 = gpu.kmeans(init, 4.0)
    assert km2["iters"] == km["iters"] and np.array_equal(km2["ids"], km["ids"])
    cb = hostlib.design_codebooks(counts, c, k, hostlib.MODE_RATIO, 0.5, hostlib.DIST_MANHATTAN)
    q = gpu.quantize(cb.tables, DEBUG_SEED, want_qv=True, want_err=True)
    sym, qv = q["symbols"], q["qv"]
    assert (qv[:, c] == 10).all() and qv[:, :c].min() >= 33 and qv[:, :c].max() < 33 + 72
    # per-line error == L1 distance between the input and the -u image, divided by the columns (integer matrix: exact)
    l1 = np.abs(rows[:, :c].astype(np.int32) - qv[:, :c].astype(np.int32)).sum(1)
    assert np.array_equal(q["line_err"], l1 / float(c))
    # the state of every symbol indexes the output alphabet of the quantizer its context selects, and decodes to qv
    kc = km["ids"].astype(np.int64)[:, None] * c + np.arange(c)[None, :]
    prev = np.concatenate([np.zeros((n, 1), np.int64), qv[:, :c - 1].astype(np.int64) - 33], axis=1)
    ctx = cb.ctx_of.reshape(-1, 72)[kc, prev]
    assert (ctx != 0xFF).all()
    qi = cb.q_off[kc].astype(np.int64) + 2 * ctx + (sym >> 7)
    assert np.array_equal(cb.smap.reshape(-1, 72)[qi, qv[:, :c].astype(np.int64) - 33], sym & 0x7F)
    assert np.array_equal(cb.qmap.reshape(-1, 72)[qi, rows[:, :c].astype(np.int64) - 33], qv[:, :c] - 33)
    with tempfile.TemporaryDirectory() as d:
        nbytes = cb.encode(d + "/big.qvz", km["ids"], sym, DEBUG_SEED)           # the coder walks the same contexts and accepts every symbol
    assert 0 < nbytes < n * c


def test_quantize_custom_distortion_matrix(gpu, oracle):
    """-D FILE: an arbitrary (non-Toeplitz, non-integer) 72x72 matrix takes the general distortion path of both walks."""
    n, c, k = 9_000, 47, 2
    rows = synth_rows(n, c, seed=314).numpy()
    ids = np.random.default_rng(2).integers(0, k, n, dtype=np.uint8)
    t = synthetic_tables(k, c, seed=21)
    t.distortion = np.random.default_rng(5).random(72 * 72) * 9.0
    _load(gpu, rows, c)
    gpu.set_clusters(k, ids)
    q = _quantize_both_paths(gpu, t, DEBUG_SEED)
    o = oracle.quantize(rows, c, ids, t, DEBUG_SEED)
    assert np.array_equal(q["symbols"], o["symbols"]) and np.array_equal(q["qv"], o["qv"])
    assert np.array_equal(q["line_err"], o["line_err"])


def test_beyond_4g_symbols(gpu, oracle):
    """Maximum sizes: one shard with more than 2^32 symbols (29.2M x 150 = 4.38e9; cfg4 on one B200 is 7x that).
    Every slot, word and draw index past 2^32 must be 64-bit clean.  Checked through properties that do not need a
    CPU pass over 4 GB: exact column sums, count totals, and the walk of the first and the last lines of the big
    shard against (a) the same lines loaded as a small shard at the same global offset and (b) the CPU oracle."""
    import torch
    from qvz_b200 import lib
    n, c, k, tail = 29_200_000, 150, 2, 4096
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2**30:
        pytest.skip("needs 60 GB of free device memory")
    dev = synth_rows(n, c, seed=99, device="cuda")
    colsum = sum(dev[lo:lo + 1_000_000, :c].sum(0, dtype=torch.int64) for lo in range(0, n, 1_000_000)).cpu().numpy()
    rows = dev.cpu().numpy()
    del dev
    torch.cuda.empty_cache()
    assert n * c > 2**32
    _load(gpu, rows, c)
    # K = 1: the centroid is the exact integer column mean
    km1 = gpu.kmeans(rows[[930_886], :c], 4.0, want_ids=False)
    assert np.array_equal(km1["means"][0], (colsum // n).astype(np.uint8))
    # K = 2: sizes, and count tables that hold every line once per column
    km = gpu.kmeans(rows[[930_886, 17_636_915], :c], 4.0)
    ids = km["ids"]
    assert km["counts"].astype(np.int64).sum() == n and np.array_equal(np.bincount(ids, minlength=k), km["counts"])
    counts = gpu.cond_counts()
    assert np.array_equal(counts.reshape(k, -1).sum(1, dtype=np.uint64), km["counts"].astype(np.uint64) * c)
    t = synthetic_tables(k, c, seed=9)
    q = gpu.quantize(t, DEBUG_SEED, want_err=True)
    small = lib.Handle(0)
    try:
        for lo in (0, n - tail):                                     # n - tail is a multiple of 4 (WELL word boundary)
            part = np.ascontiguousarray(rows[lo:lo + tail])
            small.load_rows(part, tail, c, c + 1, first_line=lo)
            small.set_clusters(k, ids[lo:lo + tail])
            qs = small.quantize(t, DEBUG_SEED, want_err=True)
            assert np.array_equal(qs["symbols"], q["symbols"][lo:lo + tail]), lo
            assert np.array_equal(qs["line_err"], q["line_err"][lo:lo + tail]), lo
            o = oracle.quantize(part, c, ids[lo:lo + tail], t, DEBUG_SEED, first_line=lo)
            assert np.array_equal(o["symbols"], qs["symbols"]), lo
            assert np.array_equal(o["line_err"], qs["line_err"]), lo
    finally:
        small.close()


@pytest.mark.parametrize("n,c", [(40_000, 150), (9_001, 301), (70_000, 23), (3, 2)])
def test_one_cluster_counting_paths(gpu, oracle, n, c):
    """K = 1 (the reference's default): the k-means column sums come out of the count table, and the table is counted
    by the lane-private byte-plane kernel (whole columns per CTA when C >= SM count, pieces of the left-over columns)
    or by the word-column kernel (alphabets above Q41, or QVZ_COUNTS_WORDS=1).  All of them must give the oracle's
    centroid, `moved` log and table -- also when the counts are asked for twice, or after new ids were installed."""
    import os
    rows = synth_rows(n, c, seed=31 + c).numpy()
    init = rows[[n // 2], :c]
    o = oracle.kmeans(rows, c, init, 4.0)
    want = oracle.cond_counts(rows, c, 1, o["ids"])
    for env in (None, "QVZ_COUNTS_WORDS"):
        if env:
            os.environ[env] = "1"
        try:
            _load(gpu, rows, c)
            r = gpu.kmeans(init, 4.0)
            assert r["iters"] == o["iters"]
            for key in ("ids", "means", "counts", "moved"):
                assert np.array_equal(r[key], o[key]), (env, key)
            assert np.array_equal(gpu.cond_counts(), want), env          # the table the k-means pass left behind
            assert np.array_equal(gpu.cond_counts(), want), env
            gpu.set_clusters(1, np.zeros(n, np.uint8))                   # new ids: counted again
            assert np.array_equal(gpu.cond_counts(), want), env
        finally:
            if env:
                del os.environ[env]
    # a wide alphabet (symbols up to 71) cannot use the 32 private copies: same results through the general kernel
    wide = rows.copy()
    wide[::7, : c] = np.minimum(wide[::7, : c].astype(np.int32) + 30, 33 + 71).astype(np.uint8)
    ow = oracle.kmeans(wide, c, wide[[0], :c], 4.0)
    _load(gpu, wide, c)
    rw = gpu.kmeans(wide[[0], :c], 4.0)
    assert np.array_equal(rw["means"], ow["means"]) and np.array_equal(rw["moved"], ow["moved"])
    assert np.array_equal(gpu.cond_counts(), oracle.cond_counts(wide, c, 1, ow["ids"]))
