"""ctypes binding of qvz_b200/csrc/libqvz_gpu.so (include/qvz_gpu.h).

There is NO CPU fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libqvz_gpu.so")
ALPHABET = 72

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)

EXPORTS = [
    "qvz_gpu_open", "qvz_gpu_close", "qvz_gpu_last_error", "qvz_gpu_stream", "qvz_gpu_get_timings",
    "qvz_gpu_reset_launch_count", "qvz_gpu_load_rows", "qvz_gpu_kmeans", "qvz_gpu_set_clusters",
    "qvz_gpu_kmeans_begin", "qvz_gpu_kmeans_assign_dev", "qvz_gpu_kmeans_update_dev", "qvz_gpu_kmeans_end",
    "qvz_gpu_kmeans_assign_host", "qvz_gpu_kmeans_update_host",
    "qvz_gpu_kmeans_update_async", "qvz_gpu_kmeans_poll", "qvz_gpu_kmeans_result", "qvz_gpu_upload_tables",
    "qvz_gpu_cond_counts", "qvz_gpu_cond_counts_dev", "qvz_gpu_cond_counts_len", "qvz_gpu_quantize",
    "qvz_gpu_prefetch_draws",
    "qvz_gpu_well_jump",
]


class QvzError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"qvz_gpu error {code}: {msg}")
        self.code = code


class FlatTablesStruct(C.Structure):
    """struct qvz_flat_tables"""
    _fields_ = [("clusters", C.c_uint32), ("columns", C.c_uint32),
                ("nctx", u32p), ("ctx_of", u8p), ("q_off", u64p), ("qratio", u8p),
                ("qmap", u8p), ("smap", u8p), ("distortion", f64p)]


class Timings(C.Structure):
    """struct qvz_gpu_timings"""
    _fields_ = [("load_h2d_ms", C.c_float), ("load_layout_ms", C.c_float), ("kmeans_ms", C.c_float),
                ("kmeans_assign_ms", C.c_float), ("cond_counts_ms", C.c_float),
                ("quantize_setup_ms", C.c_float), ("quantize_ms", C.c_float), ("quantize_draws_ms", C.c_float),
                ("quantize_d2h_ms", C.c_float),
                ("kmeans_iters", C.c_uint32), ("kernel_launches", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(verbose: bool = False) -> str:
    """Compile libqvz_gpu.so for sm_100a with the in-tree Makefile (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout, r.stderr)
    if r.returncode:
        raise RuntimeError("building libqvz_gpu.so failed")
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """dlopen the CUDA library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(the qvz front end has no CPU fallback)")
    L = C.CDLL(os.environ.get("QVZ_GPU_LIB", LIB_PATH))     # (QVZ_GPU_LIB: an experimental build of the same library)
    vp = C.c_void_p
    L.qvz_gpu_open.restype = C.c_int
    L.qvz_gpu_open.argtypes = [C.POINTER(vp), C.c_int]
    L.qvz_gpu_close.restype = None
    L.qvz_gpu_close.argtypes = [vp]
    L.qvz_gpu_last_error.restype = C.c_char_p
    L.qvz_gpu_last_error.argtypes = [vp]
    L.qvz_gpu_stream.restype = vp
    L.qvz_gpu_stream.argtypes = [vp]
    L.qvz_gpu_get_timings.restype = C.c_int
    L.qvz_gpu_get_timings.argtypes = [vp, C.POINTER(Timings)]
    L.qvz_gpu_reset_launch_count.restype = C.c_int
    L.qvz_gpu_reset_launch_count.argtypes = [vp]
    L.qvz_gpu_load_rows.restype = C.c_int
    L.qvz_gpu_load_rows.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
    L.qvz_gpu_kmeans.restype = C.c_int
    L.qvz_gpu_kmeans.argtypes = [vp, C.c_uint32, u8p, C.c_double, C.c_uint32, u8p, u8p, u32p, f64p, u32p]
    L.qvz_gpu_set_clusters.restype = C.c_int
    L.qvz_gpu_set_clusters.argtypes = [vp, C.c_uint32, u8p]
    L.qvz_gpu_kmeans_begin.restype = C.c_int
    L.qvz_gpu_kmeans_begin.argtypes = [vp, C.c_uint32, u8p]
    L.qvz_gpu_kmeans_assign_dev.restype = C.c_int
    L.qvz_gpu_kmeans_assign_dev.argtypes = [vp, vp]
    L.qvz_gpu_kmeans_update_dev.restype = C.c_int
    L.qvz_gpu_kmeans_update_dev.argtypes = [vp, vp, f64p, u32p]
    L.qvz_gpu_kmeans_assign_host.restype = C.c_int
    L.qvz_gpu_kmeans_assign_host.argtypes = [vp, i64p]
    L.qvz_gpu_kmeans_update_host.restype = C.c_int
    L.qvz_gpu_kmeans_update_host.argtypes = [vp, i64p, f64p, u32p]
    L.qvz_gpu_kmeans_end.restype = C.c_int
    L.qvz_gpu_kmeans_end.argtypes = [vp, u8p, u8p]
    L.qvz_gpu_kmeans_update_async.restype = C.c_int
    L.qvz_gpu_kmeans_update_async.argtypes = [vp, vp, C.c_double, C.c_uint32]
    L.qvz_gpu_kmeans_poll.restype = C.c_int
    L.qvz_gpu_kmeans_poll.argtypes = [vp, C.c_uint32, C.POINTER(C.c_int), u32p]
    L.qvz_gpu_kmeans_result.restype = C.c_int
    L.qvz_gpu_kmeans_result.argtypes = [vp, u32p, f64p, u32p]
    L.qvz_gpu_upload_tables.restype = C.c_int
    L.qvz_gpu_upload_tables.argtypes = [vp, C.POINTER(FlatTablesStruct)]
    L.qvz_gpu_cond_counts.restype = C.c_int
    L.qvz_gpu_cond_counts.argtypes = [vp, u32p]
    L.qvz_gpu_cond_counts_dev.restype = C.c_int
    L.qvz_gpu_cond_counts_dev.argtypes = [vp, vp]
    L.qvz_gpu_cond_counts_len.restype = C.c_uint64
    L.qvz_gpu_cond_counts_len.argtypes = [C.c_uint32, C.c_uint32]
    L.qvz_gpu_quantize.restype = C.c_int
    L.qvz_gpu_quantize.argtypes = [vp, C.POINTER(FlatTablesStruct), u32p, vp, vp, vp]
    L.qvz_gpu_prefetch_draws.restype = C.c_int
    L.qvz_gpu_prefetch_draws.argtypes = [vp, u32p]
    L.qvz_gpu_well_jump.restype = C.c_int
    L.qvz_gpu_well_jump.argtypes = [vp, u32p, C.c_uint64, u32p]
    _lib = L
    return L


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _addr(a):
    """Host address of a numpy array / torch CPU tensor / raw int (None -> NULL)."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags.c_contiguous
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())        # torch tensor


def tables_struct(t) -> FlatTablesStruct:
    """Any object with the numpy fields of `struct qvz_flat_tables` -> ctypes struct (keeps no copy)."""
    return FlatTablesStruct(int(t.clusters), int(t.columns), _p(t.nctx, u32p), _p(t.ctx_of, u8p),
                            _p(t.q_off, u64p), _p(t.qratio, u8p), _p(t.qmap, u8p), _p(t.smap, u8p),
                            _p(t.distortion, f64p))


class Handle:
    """One qvz_gpu handle = one device = one shard of lines."""

    def __init__(self, device: int = 0):
        self.L = load()
        self.h = C.c_void_p()
        rc = self.L.qvz_gpu_open(C.byref(self.h), device)
        if rc:
            msg = self.L.qvz_gpu_last_error(self.h).decode() if self.h else "cannot open CUDA device"
            if self.h:
                self.L.qvz_gpu_close(self.h)
            self.h = None
            raise QvzError(rc, msg)
        self.n_lines = self.columns = 0
        self.K = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.qvz_gpu_close(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc:
            raise QvzError(rc, self.L.qvz_gpu_last_error(self.h).decode())

    @property
    def stream(self) -> int:
        return int(self.L.qvz_gpu_stream(self.h) or 0)

    def timings(self) -> dict:
        t = Timings()
        self._check(self.L.qvz_gpu_get_timings(self.h, C.byref(t)))
        return t.as_dict()

    def reset_launch_count(self):
        self._check(self.L.qvz_gpu_reset_launch_count(self.h))

    # ---- stage calls (host buffers: numpy arrays or pinned torch tensors) ----
    def load_rows(self, rows, n_lines: int, columns: int, row_stride: int, first_line: int = 0):
        self._check(self.L.qvz_gpu_load_rows(self.h, _addr(rows), n_lines, columns, row_stride, first_line))
        self.n_lines, self.columns = n_lines, columns

    def kmeans(self, init_means: np.ndarray, threshold: float = 4.0, max_iter: int = 1000,
               want_ids: bool = True, ids_out=None):
        init_means = np.ascontiguousarray(init_means, dtype=np.uint8)
        K = init_means.shape[0]
        ids = ids_out if ids_out is not None else (np.empty(self.n_lines, np.uint8) if want_ids else None)
        means = np.zeros((K, self.columns), np.uint8)
        counts = np.zeros(K, np.uint32)
        moved = np.zeros((max_iter, K), np.float64)
        iters = C.c_uint32(0)
        ids_p = None if ids is None else C.cast(_addr(ids), u8p)
        self._check(self.L.qvz_gpu_kmeans(self.h, K, _p(init_means, u8p), float(threshold), max_iter,
                                          ids_p, _p(means, u8p), _p(counts, u32p), _p(moved, f64p),
                                          C.byref(iters)))
        self.K = K
        return dict(iters=iters.value, ids=ids, means=means, counts=counts, moved=moved[:iters.value])

    def set_clusters(self, K: int, ids: np.ndarray):
        ids = np.ascontiguousarray(ids, dtype=np.uint8)
        assert ids.shape[0] == self.n_lines
        self._check(self.L.qvz_gpu_set_clusters(self.h, K, _p(ids, u8p)))
        self.K = K

    def cond_counts(self, want: bool = True):
        out = np.empty((self.K, 1 + ALPHABET * (self.columns - 1), ALPHABET), np.uint32) if want else None
        self._check(self.L.qvz_gpu_cond_counts(self.h, _p(out, u32p)))
        return out

    def quantize(self, tables, seed, want_symbols=True, want_qv=False, want_err=False,
                 symbols_out=None, qv_out=None, err_out=None):
        seed = np.ascontiguousarray(seed, dtype=np.uint32)
        n, c = self.n_lines, self.columns
        sym = symbols_out if symbols_out is not None else (np.empty((n, c), np.uint8) if want_symbols else None)
        qv = qv_out if qv_out is not None else (np.empty((n, c + 1), np.uint8) if want_qv else None)
        err = err_out if err_out is not None else (np.empty(n, np.float64) if want_err else None)
        if tables is None:                     # walk with the tables installed by upload_tables()
            tp = None
        else:
            st = tables if isinstance(tables, FlatTablesStruct) else tables_struct(tables)
            tp = C.byref(st)
        self._check(self.L.qvz_gpu_quantize(self.h, tp, _p(seed, u32p), _addr(sym), _addr(qv), _addr(err)))
        return dict(symbols=sym, qv=qv, line_err=err)

    def upload_tables(self, tables):
        """Install the quantizer tables on the device (stage-3 input); quantize(None, seed) then walks with them."""
        st = tables if isinstance(tables, FlatTablesStruct) else tables_struct(tables)
        self._check(self.L.qvz_gpu_upload_tables(self.h, C.byref(st)))

    def prefetch_draws(self, seed):
        seed = np.ascontiguousarray(seed, dtype=np.uint32)
        self._check(self.L.qvz_gpu_prefetch_draws(self.h, _p(seed, u32p)))

    def well_jump(self, seed, words: int) -> np.ndarray:
        seed = np.ascontiguousarray(seed, dtype=np.uint32)
        out = np.zeros(32, np.uint32)
        self._check(self.L.qvz_gpu_well_jump(self.h, _p(seed, u32p), words, _p(out, u32p)))
        return out

    # ---- stepping interface (device pointers; used by qvz_b200.dist) ----
    def kmeans_begin(self, init_means: np.ndarray):
        init_means = np.ascontiguousarray(init_means, dtype=np.uint8)
        self._check(self.L.qvz_gpu_kmeans_begin(self.h, init_means.shape[0], _p(init_means, u8p)))
        self.K = init_means.shape[0]

    def kmeans_assign_dev(self, sums_ptr: int):
        self._check(self.L.qvz_gpu_kmeans_assign_dev(self.h, C.c_void_p(sums_ptr)))

    def kmeans_update_dev(self, sums_ptr: int):
        moved = np.zeros(self.K, np.float64)
        counts = np.zeros(self.K, np.uint32)
        self._check(self.L.qvz_gpu_kmeans_update_dev(self.h, C.c_void_p(sums_ptr), _p(moved, f64p), _p(counts, u32p)))
        return moved, counts

    def kmeans_update_async(self, sums_ptr: int, threshold: float, max_iter: int):
        """Recentre and decide on the device whether the loop goes on; does not wait for the GPU."""
        self._check(self.L.qvz_gpu_kmeans_update_async(self.h, C.c_void_p(sums_ptr), float(threshold), int(max_iter)))

    def kmeans_poll(self, idx: int):
        """(done, iterations completed) after the idx-th update_async of this run."""
        done, iters = C.c_int(0), C.c_uint32(0)
        self._check(self.L.qvz_gpu_kmeans_poll(self.h, idx, C.byref(done), C.byref(iters)))
        return bool(done.value), int(iters.value)

    def kmeans_result(self, max_rows: int = 1000):
        """(iterations, moved log [iterations, K], line counts [K]) of the finished run."""
        iters = C.c_uint32(0)
        moved = np.zeros((max_rows, self.K), np.float64)
        counts = np.zeros(self.K, np.uint32)
        self._check(self.L.qvz_gpu_kmeans_result(self.h, C.byref(iters), _p(moved, f64p), _p(counts, u32p)))
        return int(iters.value), moved[:min(iters.value, max_rows)], counts

    def kmeans_end(self, want_ids=True):
        ids = np.empty(self.n_lines, np.uint8) if want_ids else None
        means = np.zeros((self.K, self.columns), np.uint8)
        self._check(self.L.qvz_gpu_kmeans_end(self.h, _p(ids, u8p), _p(means, u8p)))
        return ids, means

    def cond_counts_dev(self, counts_ptr: int):
        self._check(self.L.qvz_gpu_cond_counts_dev(self.h, C.c_void_p(counts_ptr)))

    def cond_counts_len(self) -> int:
        return int(self.L.qvz_gpu_cond_counts_len(self.K, self.columns))
