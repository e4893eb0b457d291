"""ctypes binding of qvz_b200/host/libqvz_host.so (include/qvz_host.h): codebook design from the GPU's conditional
counts, flat tables for the GPU quantize walk, and the .qvz writer (codebooks + seed + arithmetic coder)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .lib import FlatTablesStruct

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "host", "libqvz_host.so")
ALPHABET = 72
MODE_RATIO, MODE_FIXED = 0, 1
DIST_MANHATTAN, DIST_MSE, DIST_LORENTZ = 1, 2, 3

EXPORTS = ["qvz_host_distortion", "qvz_host_distortion_file", "qvz_host_design", "qvz_host_free", "qvz_host_tables",
           "qvz_host_codebook_bytes", "qvz_host_write_codebooks", "qvz_host_encode", "qvz_host_decode"]

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
f64p = C.POINTER(C.c_double)
_lib = None


def build(verbose: bool = False) -> str:
    r = subprocess.run(["make", "-C", os.path.join(HERE, "host"), "libqvz_host.so"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout, r.stderr)
    if r.returncode:
        raise RuntimeError("building libqvz_host.so failed")
    return LIB_PATH


def available() -> bool:
    return os.path.exists(LIB_PATH)


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not available():
            build()
        L = C.CDLL(LIB_PATH)
        L.qvz_host_distortion.restype = C.c_int
        L.qvz_host_distortion.argtypes = [C.c_int, f64p]
        L.qvz_host_distortion_file.restype = C.c_int
        L.qvz_host_distortion_file.argtypes = [C.c_char_p, f64p]
        L.qvz_host_design.restype = C.c_void_p
        L.qvz_host_design.argtypes = [u32p, C.c_uint32, C.c_uint32, C.c_int, C.c_double, f64p, C.c_int]
        L.qvz_host_free.restype = None
        L.qvz_host_free.argtypes = [C.c_void_p]
        L.qvz_host_tables.restype = C.c_int
        L.qvz_host_tables.argtypes = [C.c_void_p, C.POINTER(FlatTablesStruct)]
        L.qvz_host_codebook_bytes.restype = C.c_uint64
        L.qvz_host_codebook_bytes.argtypes = [C.c_void_p]
        L.qvz_host_write_codebooks.restype = C.c_int
        L.qvz_host_write_codebooks.argtypes = [C.c_void_p, C.c_uint64, u8p]
        L.qvz_host_encode.restype = C.c_int
        L.qvz_host_encode.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, u8p, u8p, u32p, C.POINTER(C.c_uint64)]
        L.qvz_host_decode.restype = C.c_int
        L.qvz_host_decode.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def distortion_matrix(kind: int) -> np.ndarray:
    out = np.zeros(ALPHABET * ALPHABET, np.float64)
    if load().qvz_host_distortion(kind, out.ctypes.data_as(f64p)):
        raise ValueError(f"unknown distortion type {kind}")
    return out


class Codebooks:
    """Owns a qvz_codebooks handle; `.tables` is a struct qvz_flat_tables view usable with lib.Handle.quantize,
    with numpy views of the same memory as attributes (nctx, ctx_of, q_off, qratio, qmap, smap, distortion)."""

    def __init__(self, counts: np.ndarray, columns: int, clusters: int, mode: int, target: float, distortion, threads: int = 0):
        L = load()
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        assert counts.size == clusters * (1 + ALPHABET * (columns - 1)) * ALPHABET
        if not isinstance(distortion, np.ndarray):
            distortion = distortion_matrix(int(distortion))
        d = np.ascontiguousarray(distortion, dtype=np.float64)
        self.L = L
        self.h = L.qvz_host_design(counts.ctypes.data_as(u32p), clusters, columns, mode, float(target), d.ctypes.data_as(f64p), threads)
        if not self.h:
            raise ValueError("qvz_host_design rejected its arguments")
        self.clusters, self.columns = clusters, columns
        self.tables = FlatTablesStruct()
        L.qvz_host_tables(self.h, C.byref(self.tables))
        t = self.tables
        kc = clusters * columns
        self.nctx = np.ctypeslib.as_array(t.nctx, (kc,))
        self.q_off = np.ctypeslib.as_array(t.q_off, (kc,))
        nq = int(self.q_off[-1] + 2 * self.nctx[-1])
        self.ctx_of = np.ctypeslib.as_array(t.ctx_of, (kc * ALPHABET,))
        self.qratio = np.ctypeslib.as_array(t.qratio, (nq // 2,))
        self.qmap = np.ctypeslib.as_array(t.qmap, (nq * ALPHABET,))
        self.smap = np.ctypeslib.as_array(t.smap, (nq * ALPHABET,))
        self.distortion = np.ctypeslib.as_array(t.distortion, (ALPHABET * ALPHABET,))

    def codebook_bytes(self, n_lines: int) -> np.ndarray:
        out = np.zeros(int(self.L.qvz_host_codebook_bytes(self.h)), np.uint8)
        if self.L.qvz_host_write_codebooks(self.h, n_lines, out.ctypes.data_as(u8p)):
            raise RuntimeError("qvz_host_write_codebooks failed")
        return out

    def encode(self, path: str, cluster_ids: np.ndarray, symbols: np.ndarray, seed: np.ndarray) -> int:
        ids = np.ascontiguousarray(cluster_ids, dtype=np.uint8)
        sym = np.ascontiguousarray(symbols, dtype=np.uint8)
        seed = np.ascontiguousarray(seed, dtype=np.uint32)
        n = ids.shape[0]
        assert sym.size == n * self.columns
        written = C.c_uint64(0)
        rc = self.L.qvz_host_encode(self.h, path.encode(), n, ids.ctypes.data_as(u8p), sym.ctypes.data_as(u8p),
                                    seed.ctypes.data_as(u32p), C.byref(written))
        if rc:
            raise RuntimeError(f"qvz_host_encode failed ({rc})")
        return written.value

    def close(self):
        if getattr(self, "h", None):
            self.L.qvz_host_free(self.h)
            self.h = None

    __del__ = close


def decode_file(src: str, dst: str) -> int:
    """qvz -x: returns the number of lines written."""
    n = C.c_uint64(0)
    rc = load().qvz_host_decode(src.encode(), dst.encode(), C.byref(n))
    if rc:
        raise RuntimeError(f"qvz_host_decode failed ({rc})")
    return n.value


def design_codebooks(counts, columns, clusters, mode, target, distortion, threads: int = 0) -> Codebooks:
    return Codebooks(counts, columns, clusters, mode, target, distortion, threads)
