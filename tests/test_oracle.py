"""The CPU oracle (oracle/qvz_oracle.c) against (1) the survey's known-answer vectors, (2) the golden
fixtures generated from the reference, (3) the compiled reference itself on fresh seeded inputs."""
import numpy as np
import pytest

from oracle.bindings import (DEBUG_SEED, DIST_MANHATTAN, DIST_MSE, MODE_FIXED, MODE_RATIO,
                             kmeans_init_lines)
from qvz_b200.synth import synth_rows

# vectors that depend only on reference code
WELL_FIRST_WORDS = [0x82fd4280, 0x815ebfd5, 0x5303e800, 0x2e7c4815, 0x8473bc42, 0xfa09693d, 0xf81bdc3d, 0x0cf612f5]
WELL_FIRST_DRAWS = [0, 5, 117, 23, 85, 127, 122, 10, 0, 80, 15, 24]
WELL_WORD_1M = 0x90e10060
GLIBC_RAND_SEED1 = [1804289383, 846930886, 1681692777, 1714636915, 1957747793, 424238335, 719885386,
                    1649760492, 596516649, 1189641421]


def test_well_known_answers(oracle):
    assert list(oracle.well_words(DEBUG_SEED, 0, 8)) == WELL_FIRST_WORDS
    assert list(oracle.well_draws(DEBUG_SEED, 12)) == WELL_FIRST_DRAWS
    assert int(oracle.well_words(DEBUG_SEED, 999_999, 1)[0]) == WELL_WORD_1M
    # after 1e6 steps n is back at 0, so the rotated frame equals the raw state and u[0] is the last output
    st = oracle.well_state_after(DEBUG_SEED, 1_000_000)
    assert int(st[0]) == WELL_WORD_1M


def test_well_draw_index_formula(oracle):
    # draw d = (word[d >> 2] >> 7*(d & 3)) & 127   
    words = oracle.well_words(DEBUG_SEED, 0, 300)
    draws = oracle.well_draws(DEBUG_SEED, 1200)
    d = np.arange(1200)
    assert np.array_equal(draws, ((words[d >> 2] >> (7 * (d & 3))) & 127).astype(np.uint8))


def test_well_vs_reference(oracle, ref):
    rng = np.random.default_rng(5)
    seed = rng.integers(0, 2**31, 32, dtype=np.uint32)      # rand() values are < 2^31 (qv_stream.c:80)
    assert np.array_equal(oracle.well_words(seed, 12345, 4000), ref.well_words(seed, 12345, 4000))
    assert np.array_equal(oracle.well_draws(seed, 9000), ref.well_draws(seed, 9000))
    assert ref.rand_stream(10) == GLIBC_RAND_SEED1


def test_golden_kmeans(oracle, golden):
    g = golden
    c = g["columns"]
    init = g["rows"][g["picks"].astype(np.int64), :c]
    o = oracle.kmeans(g["rows"], c, init, float(g["threshold"]))
    assert o["iters"] == int(g["iters"])
    assert np.array_equal(o["ids"], g["ids"])
    assert np.array_equal(o["means"], g["means"])
    assert np.array_equal(o["counts"], g["kcounts"])
    assert np.array_equal(o["moved"], g["moved"])


def test_golden_cond_counts(oracle, golden):
    g = golden
    o = oracle.cond_counts(g["rows"], g["columns"], g["clusters"], g["ids"])
    assert np.array_equal(o, g["cond_counts"])
    assert int(o.sum()) == g["rows"].shape[0] * g["columns"]


def test_golden_quantize(oracle, golden):
    g = golden
    o = oracle.quantize(g["rows"], g["columns"], g["ids"], g["tables"], DEBUG_SEED)
    assert np.array_equal(o["symbols"], g["symbols"])
    assert np.array_equal(o["qv"], g["qv"])
    assert np.array_equal(o["line_err"], g["line_err"])          # bit-exact doubles
    assert o["distortion"] == float(g["distortion"])


def test_quantize_sharded_draw_offset(oracle, golden):
    # a shard starting at line L0 must use draws L0*C.. : quantizing [L0:] alone equals the tail of the whole
    g = golden
    L0 = 1000
    o = oracle.quantize(np.ascontiguousarray(g["rows"][L0:]), g["columns"], g["ids"][L0:], g["tables"],
                        DEBUG_SEED, first_line=L0)
    assert np.array_equal(o["symbols"], g["symbols"][L0:])
    assert np.array_equal(o["line_err"], g["line_err"][L0:])


@pytest.mark.parametrize("n,c,k,thr", [(5000, 37, 4, 4.0), (1200, 9, 2, 0.0), (2000, 150, 3, 10.0), (7, 5, 1, 4.0)])
def test_kmeans_and_counts_vs_reference(oracle, ref, n, c, k, thr):
    rows = synth_rows(n, c, seed=100 + n).numpy()
    picks = kmeans_init_lines(n, k, ref.rand_stream(2 * k))
    s = ref.session(rows, c, k, threshold=thr)
    r = s.kmeans(picks)
    o = oracle.kmeans(rows, c, rows[picks, :c], thr)
    assert o["iters"] == r["iters"]
    for key in ("ids", "means", "counts", "moved"):
        assert np.array_equal(o[key], r[key]), key
    rc, totals = s.stats()
    oc = oracle.cond_counts(rows, c, k, o["ids"])
    assert np.array_equal(oc, rc)
    assert np.array_equal(oc.sum(-1), totals)


def test_kmeans_reference_driver_equals_stepped_loop(ref):
    # the harness loop (used to log per-iteration values) reproduces do_kmeans_clustering itself
    n, c, k = 3000, 20, 3
    rows = synth_rows(n, c, seed=42).numpy()
    ids_driver = ref.session(rows, c, k).kmeans_reference_driver()
    picks = kmeans_init_lines(n, k, ref.rand_stream(2 * k))
    assert np.array_equal(ref.session(rows, c, k).kmeans(picks)["ids"], ids_driver)


@pytest.mark.parametrize("mode,ratio,dist,k", [(MODE_RATIO, 0.3, DIST_MSE, 2), (MODE_FIXED, 1.5, DIST_MANHATTAN, 1)])
def test_quantize_vs_reference(oracle, ref, mode, ratio, dist, k):
    n, c = 1500, 10
    rows = synth_rows(n, c, seed=77).numpy()
    picks = kmeans_init_lines(n, k, ref.rand_stream(2 * k))
    s = ref.session(rows, c, k, mode=mode, ratio=ratio, distortion=dist)
    ids = s.kmeans(picks)["ids"]
    t = s.tables()
    seed = np.random.default_rng(1).integers(0, 2**31, 32, dtype=np.uint32)
    r = s.quantize(seed)
    o = oracle.quantize(rows, c, ids, t, seed)
    assert np.array_equal(o["symbols"], r["symbols"])
    assert np.array_equal(o["qv"], r["qv"])
    assert np.array_equal(o["line_err"], r["line_err"])
    assert o["distortion"] == r["distortion"]


def test_empty_cluster_is_reported(oracle):
    # two identical initial centroids: the second cluster never wins a tie (strict '<') -> empty -> the
    # reference divides by zero (src/cluster.c:113); the oracle reports -1 instead
    rows = synth_rows(100, 8, seed=3).numpy()
    init = np.stack([rows[0, :8], rows[0, :8]])
    assert oracle.kmeans(rows, 8, init)["iters"] == -1
