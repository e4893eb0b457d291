"""ctypes bindings for the CPU oracle and the compiled reference -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (qvz_b200/) never does.

  Oracle  -> oracle/liboracle.so      (oracle/qvz_oracle.c, CPU restatement)
  Ref     -> oracle/_ref/libqvzref.so (unmodified reference sources + oracle/ref_harness.c)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ALPHABET = 72

MODE_RATIO, MODE_FIXED = 0, 1                       # include/codebook.h:22-23
DIST_MANHATTAN, DIST_MSE, DIST_LORENTZ = 1, 2, 3    # include/distortion.h:7-9

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def build(force: bool = False) -> None:
    """Compile liboracle.so (always possible) and, when /root/reference exists, oracle/_ref/."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")) or \
            os.path.getmtime(os.path.join(HERE, "liboracle.so")) < os.path.getmtime(os.path.join(HERE, "qvz_oracle.c")):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)
    if os.path.exists("/root/reference/src/main.c"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


class FlatTablesStruct(C.Structure):
    """struct qvz_flat_tables (include/qvz_gpu.h)."""
    _fields_ = [("clusters", C.c_uint32), ("columns", C.c_uint32),
                ("nctx", u32p), ("ctx_of", u8p), ("q_off", u64p), ("qratio", u8p),
                ("qmap", u8p), ("smap", u8p), ("distortion", f64p)]


@dataclass
class FlatTables:
    """numpy-owned flat mirror of cond_quantizer_list_t for all clusters."""
    clusters: int
    columns: int
    nctx: np.ndarray        # uint32 [K*C]
    ctx_of: np.ndarray      # uint8  [K*C*72]
    q_off: np.ndarray       # uint64 [K*C]
    qratio: np.ndarray      # uint8  [nq/2]
    qmap: np.ndarray        # uint8  [nq*72]
    smap: np.ndarray        # uint8  [nq*72]
    distortion: np.ndarray  # float64 [72*72]

    def as_struct(self) -> FlatTablesStruct:
        return FlatTablesStruct(self.clusters, self.columns, _p(self.nctx, u32p), _p(self.ctx_of, u8p),
                                _p(self.q_off, u64p), _p(self.qratio, u8p), _p(self.qmap, u8p),
                                _p(self.smap, u8p), _p(self.distortion, f64p))

    def save(self, path: str) -> None:
        np.savez_compressed(path, clusters=self.clusters, columns=self.columns, nctx=self.nctx,
                            ctx_of=self.ctx_of, q_off=self.q_off, qratio=self.qratio, qmap=self.qmap,
                            smap=self.smap, distortion=self.distortion)

    @staticmethod
    def load(path: str) -> "FlatTables":
        z = np.load(path)
        return FlatTables(int(z["clusters"]), int(z["columns"]), z["nctx"], z["ctx_of"], z["q_off"],
                          z["qratio"], z["qmap"], z["smap"], z["distortion"])


def _rows_2d(rows: np.ndarray) -> np.ndarray:
    assert rows.dtype == np.uint8 and rows.ndim == 2 and rows.flags.c_contiguous
    return rows


class Oracle:
    """oracle/liboracle.so -- rows are uint8 [N, stride] arrays, stride >= columns."""

    def __init__(self) -> None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.oracle_kmeans.restype = C.c_int32
        L.oracle_kmeans.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u8p, C.c_double,
                                    C.c_uint32, u8p, u8p, u32p, f64p]
        L.oracle_cond_counts.restype = None
        L.oracle_cond_counts.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u8p, u32p]
        L.oracle_quantize.restype = C.c_double
        L.oracle_quantize.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, u8p,
                                      C.POINTER(FlatTablesStruct), u32p, u8p, u8p, f64p]
        L.oracle_well_words.restype = None
        L.oracle_well_words.argtypes = [u32p, C.c_uint64, C.c_uint64, u32p]
        L.oracle_well_draws.restype = None
        L.oracle_well_draws.argtypes = [u32p, C.c_uint64, u8p]
        L.oracle_well_state_after.restype = None
        L.oracle_well_state_after.argtypes = [u32p, C.c_uint64, u32p]

    def kmeans(self, rows, columns, init_means, threshold=4.0, max_iter=1000):
        rows = _rows_2d(rows)
        n, stride = rows.shape
        init_means = np.ascontiguousarray(init_means, dtype=np.uint8)
        K = init_means.shape[0]
        ids = np.zeros(n, np.uint8)
        means = np.zeros((K, columns), np.uint8)
        counts = np.zeros(K, np.uint32)
        moved = np.zeros((max_iter, K), np.float64)
        it = self.lib.oracle_kmeans(_p(rows, u8p), n, columns, stride, K, _p(init_means, u8p),
                                    float(threshold), max_iter, _p(ids, u8p), _p(means, u8p),
                                    _p(counts, u32p), _p(moved, f64p))
        return dict(iters=it, ids=ids, means=means, counts=counts, moved=moved[:max(it, 0)])

    def cond_counts(self, rows, columns, K, ids):
        rows = _rows_2d(rows)
        n, stride = rows.shape
        ids = np.ascontiguousarray(ids, dtype=np.uint8)
        out = np.zeros((K, 1 + ALPHABET * (columns - 1), ALPHABET), np.uint32)
        self.lib.oracle_cond_counts(_p(rows, u8p), n, columns, stride, K, _p(ids, u8p), _p(out, u32p))
        return out

    def quantize(self, rows, columns, ids, tables: FlatTables, seed, first_line=0, want_qv=True,
                 want_err=True):
        rows = _rows_2d(rows)
        n, stride = rows.shape
        ids = np.ascontiguousarray(ids, dtype=np.uint8)
        seed = np.ascontiguousarray(seed, dtype=np.uint32)
        sym = np.zeros((n, columns), np.uint8)
        qv = np.zeros((n, columns + 1), np.uint8) if want_qv else None
        err = np.zeros(n, np.float64) if want_err else None
        st = tables.as_struct()
        d = self.lib.oracle_quantize(_p(rows, u8p), n, columns, stride, first_line, _p(ids, u8p),
                                     C.byref(st), _p(seed, u32p), _p(sym, u8p), _p(qv, u8p), _p(err, f64p))
        return dict(distortion=d, symbols=sym, qv=qv, line_err=err)

    def well_words(self, seed, skip, count):
        seed = np.ascontiguousarray(seed, dtype=np.uint32)
        out = np.zeros(count, np.uint32)
        self.lib.oracle_well_words(_p(seed, u32p), skip, count, _p(out, u32p))
        return out

    def well_draws(self, seed, count):
        seed = np.ascontiguousarray(seed, dtype=np.uint32)
        out = np.zeros(count, np.uint8)
        self.lib.oracle_well_draws(_p(seed, u32p), count, _p(out, u8p))
        return out

    def well_state_after(self, seed, words):
        seed = np.ascontiguousarray(seed, dtype=np.uint32)
        out = np.zeros(32, np.uint32)
        self.lib.oracle_well_state_after(_p(seed, u32p), words, _p(out, u32p))
        return out


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libqvzref.so"))


class Ref:
    """oracle/_ref/libqvzref.so -- the unmodified reference behind oracle/ref_harness.c."""

    def __init__(self) -> None:
        path = os.path.join(HERE, "_ref", "libqvzref.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_double, C.c_int]
        L.ref_kmeans_reference_driver.restype = None
        L.ref_kmeans_reference_driver.argtypes = [C.c_void_p, u8p]
        L.ref_kmeans.restype = C.c_uint32
        L.ref_kmeans.argtypes = [C.c_void_p, u64p, C.c_uint32, u8p, u8p, u32p, f64p]
        L.ref_set_clusters.restype = None
        L.ref_set_clusters.argtypes = [C.c_void_p, u8p]
        L.ref_stats.restype = None
        L.ref_stats.argtypes = [C.c_void_p, u32p, u32p]
        L.ref_codebooks.restype = None
        L.ref_codebooks.argtypes = [C.c_void_p]
        L.ref_tables_count.restype = C.c_uint64
        L.ref_tables_count.argtypes = [C.c_void_p]
        L.ref_tables_export.restype = None
        L.ref_tables_export.argtypes = [C.c_void_p, u32p, u8p, u64p, u8p, u8p, u8p, f64p]
        L.ref_quantize.restype = C.c_double
        L.ref_quantize.argtypes = [C.c_void_p, u32p, u8p, u8p, f64p]
        L.ref_well_words.restype = None
        L.ref_well_words.argtypes = [u32p, C.c_uint64, C.c_uint64, u32p]
        L.ref_well_draws.restype = None
        L.ref_well_draws.argtypes = [u32p, C.c_uint64, u8p]
        L.ref_encode_file.restype = None
        L.ref_encode_file.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_double, C.c_int,
                                      C.c_double, C.c_int]
        L.ref_decode_file.restype = None
        L.ref_decode_file.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_rand_stream.restype = None
        L.ref_rand_stream.argtypes = [C.c_uint32, C.POINTER(C.c_int32)]

    class Session:
        """One quality_file_t over a [N, C+1] '\\n'-terminated row image."""

        def __init__(self, ref: "Ref", rows: np.ndarray, columns: int, clusters: int, threshold=4.0,
                     mode=MODE_RATIO, ratio=0.5, distortion=DIST_MSE):
            rows = _rows_2d(rows)
            assert rows.shape[1] == columns + 1, "the reference needs the file image (stride C+1)"
            self.ref, self.rows, self.columns, self.K = ref, rows, columns, clusters
            self.n = rows.shape[0]
            self.h = ref.lib.ref_create(_p(rows, u8p), self.n, columns, clusters, float(threshold), mode,
                                        float(ratio), distortion)
            assert self.h

        def kmeans_reference_driver(self):
            ids = np.zeros(self.n, np.uint8)
            self.ref.lib.ref_kmeans_reference_driver(self.h, _p(ids, u8p))
            return ids

        def kmeans(self, init_lines, max_iter=1000):
            init_lines = np.ascontiguousarray(init_lines, dtype=np.uint64)
            ids = np.zeros(self.n, np.uint8)
            means = np.zeros((self.K, self.columns), np.uint8)
            counts = np.zeros(self.K, np.uint32)
            moved = np.zeros((max_iter, self.K), np.float64)
            it = self.ref.lib.ref_kmeans(self.h, _p(init_lines, u64p), max_iter, _p(ids, u8p), _p(means, u8p),
                                         _p(counts, u32p), _p(moved, f64p))
            return dict(iters=it, ids=ids, means=means, counts=counts, moved=moved[:it])

        def set_clusters(self, ids):
            ids = np.ascontiguousarray(ids, dtype=np.uint8)
            self.ref.lib.ref_set_clusters(self.h, _p(ids, u8p))

        def stats(self):
            rows = 1 + ALPHABET * (self.columns - 1)
            counts = np.zeros((self.K, rows, ALPHABET), np.uint32)
            totals = np.zeros((self.K, rows), np.uint32)
            self.ref.lib.ref_stats(self.h, _p(counts, u32p), _p(totals, u32p))
            return counts, totals

        def tables(self) -> FlatTables:
            self.ref.lib.ref_codebooks(self.h)
            nq = int(self.ref.lib.ref_tables_count(self.h))
            KC = self.K * self.columns
            t = FlatTables(self.K, self.columns, np.zeros(KC, np.uint32), np.zeros(KC * ALPHABET, np.uint8),
                           np.zeros(KC, np.uint64), np.zeros(nq // 2, np.uint8),
                           np.zeros(nq * ALPHABET, np.uint8), np.zeros(nq * ALPHABET, np.uint8),
                           np.zeros(ALPHABET * ALPHABET, np.float64))
            self.ref.lib.ref_tables_export(self.h, _p(t.nctx, u32p), _p(t.ctx_of, u8p), _p(t.q_off, u64p),
                                           _p(t.qratio, u8p), _p(t.qmap, u8p), _p(t.smap, u8p),
                                           _p(t.distortion, f64p))
            return t

        def quantize(self, seed, want_qv=True, want_err=True):
            self.ref.lib.ref_codebooks(self.h)
            seed = np.ascontiguousarray(seed, dtype=np.uint32)
            sym = np.zeros((self.n, self.columns), np.uint8)
            qv = np.zeros((self.n, self.columns + 1), np.uint8) if want_qv else None
            err = np.zeros(self.n, np.float64) if want_err else None
            d = self.ref.lib.ref_quantize(self.h, _p(seed, u32p), _p(sym, u8p), _p(qv, u8p), _p(err, f64p))
            return dict(distortion=d, symbols=sym, qv=qv, line_err=err)

    def session(self, rows, columns, clusters, **kw) -> "Ref.Session":
        return Ref.Session(self, rows, columns, clusters, **kw)

    def well_words(self, seed, skip, count):
        seed = np.ascontiguousarray(seed, dtype=np.uint32)
        out = np.zeros(count, np.uint32)
        self.lib.ref_well_words(_p(seed, u32p), skip, count, _p(out, u32p))
        return out

    def well_draws(self, seed, count):
        seed = np.ascontiguousarray(seed, dtype=np.uint32)
        out = np.zeros(count, np.uint8)
        self.lib.ref_well_draws(_p(seed, u32p), count, _p(out, u8p))
        return out

    def rand_stream(self, count):
        out = (C.c_int32 * count)()
        self.lib.ref_rand_stream(count, out)
        return list(out)

    def encode_file(self, src, dst, ufile=None, clusters=1, threshold=4.0, mode=MODE_RATIO, ratio=0.5,
                    distortion=DIST_MSE):
        self.lib.ref_encode_file(src.encode(), dst.encode(), ufile.encode() if ufile else None, clusters,
                                 float(threshold), mode, float(ratio), distortion)
        C.CDLL(None).fflush(None)    # encode() never closes the -u file (src/main.c:80-95)

    def decode_file(self, src, dst):
        self.lib.ref_decode_file(src.encode(), dst.encode())


DEBUG_SEED = np.full(32, 0x55555555, np.uint32)     # src/qv_stream.c:82


def kmeans_init_lines(n_lines: int, K: int, rand_stream) -> list[int]:
    """Global line indices picked by initialize_kmeans_clustering (src/cluster.c:199-200) given the
    libc rand() values it would consume (2 per cluster), for blocks of MAX_LINES_PER_BLOCK = 1e6."""
    MAXB = 1_000_000
    block_count = (n_lines + MAXB - 1) // MAXB
    picks = []
    for j in range(K):
        b = rand_stream[2 * j] % block_count
        cnt = MAXB if b < block_count - 1 or n_lines % MAXB == 0 else n_lines % MAXB
        picks.append(b * MAXB + rand_stream[2 * j + 1] % cnt)
    return picks
