"""Test helpers: fast synthetic (but structurally valid) quantizer tables for arbitrary shapes."""
import numpy as np

from oracle.bindings import FlatTables

ALPHABET = 72


def synthetic_tables(K: int, C: int, seed: int = 0, nsym: int = 42, dist: str = "L", mixing: bool = True) -> FlatTables:
    """Valid `struct qvz_flat_tables` without running the (slow) codebook design: every column has
    all 72 values as contexts (so no context can be missing), each context gets a coarse (lo) and a
    finer (hi) staircase quantizer with random steps/offsets and a random mixing ratio in [0,128]
    (mixing=False: only the ratios 0 and 128 -- no WELL draw can change a symbol)."""
    rng = np.random.default_rng(seed)
    KC = K * C
    nctx = np.full(KC, ALPHABET, np.uint32)
    nctx[::C] = 1                                  # column 0 has the single context {0}
    ctx_of = np.tile(np.arange(ALPHABET, dtype=np.uint8), KC)
    for k in range(K):
        base = (k * C) * ALPHABET
        ctx_of[base:base + ALPHABET] = 0xFF
        ctx_of[base] = 0
    q_off = np.zeros(KC, np.uint64)
    q_off[1:] = np.cumsum(2 * nctx.astype(np.uint64))[:-1]
    nq = int(2 * nctx.sum())
    qratio = rng.integers(0, 129, nq // 2, dtype=np.uint8)
    if not mixing:                                 # every context always takes the same one of its two quantizers (what -f 1.0 designs)
        qratio = np.where(rng.integers(0, 2, nq // 2) == 1, 128, 0).astype(np.uint8)
    x = np.arange(ALPHABET)
    step = rng.integers(1, 9, nq)
    step[1::2] = np.maximum(1, step[0::2] - rng.integers(0, 3, nq // 2))      # hi is at least as fine as lo
    off = rng.integers(0, 4, nq)
    qm = np.minimum(((x[None, :] + off[:, None]) // step[:, None]) * step[:, None] + step[:, None] // 2, ALPHABET - 1)
    qm = np.maximum(qm - off[:, None], 0).astype(np.uint8)
    qm = np.maximum.accumulate(qm, axis=1)         # monotone like a Lloyd-Max map
    smap = np.full((nq, ALPHABET), 0xFF, np.uint8)
    for i in range(nq):                            # state = index in the sorted output alphabet
        u = np.unique(qm[i])
        smap[i, u] = np.arange(u.size, dtype=np.uint8)
    d = np.abs(x[:, None] - x[None, :]).astype(np.float64)
    D = {"L": np.log2(1.0 + d), "M": d * d, "A": d}[dist]
    return FlatTables(K, C, nctx, ctx_of, q_off, qratio, qm.reshape(-1).copy(), smap.reshape(-1).copy(),
                      np.ascontiguousarray(D.T).reshape(-1).copy())
