"""World-size-2 (and 3) runs of the read-sharded front end (qvz_b200/dist.py) over gloo on the CPU.

The host-side logic under test is the product's: shard boundaries, the broadcast of the initial centroids,
the all-reduce of the integer sums inside the k-means loop, the convergence test and the count-table
all-reduce.  The per-shard stage calls, which the product sends to the CUDA library, are served here by a
stand-in handle that computes the same integer quantities with numpy / the CPU oracle, so that the
distributed result can be compared with the single-process oracle on the whole file.  (The GPU version of
this comparison is tests/test_gpu_parity.py::test_sharded_equals_whole.)"""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from qvz_b200.dist import ShardedFrontEnd, kmeans_pick_lines, shard_bounds  # noqa: E402
from qvz_b200.synth import synth_rows  # noqa: E402


def _view(ptr, n, dtype):
    ct = {np.int64: ctypes.c_int64, np.int32: ctypes.c_int32}[dtype]
    return np.ctypeslib.as_array((ct * n).from_address(ptr))


class StandInHandle:
    """Same stepping interface as qvz_b200.lib.Handle, on host memory (tests only)."""
    stream = 0

    def __init__(self):
        from oracle.bindings import Oracle
        self.O = Oracle()

    def load_rows(self, rows, n_lines, columns, row_stride, first_line=0):
        self.rows, self.n_lines, self.columns, self.first_line = rows, n_lines, columns, first_line

    def kmeans_begin(self, init):
        self.means = init.copy()
        self.K = init.shape[0]
        self.ids = np.zeros(self.n_lines, np.uint8)
        self.done, self.iters, self.log, self.states, self.last_counts = False, 0, [], [], None

    def kmeans_assign_dev(self, ptr):
        if self.done:                              # enqueued speculatively after the run ended: the kernels return at once
            return
        K, C = self.K, self.columns
        x = self.rows[:, :C].astype(np.int64)
        d = ((x[:, None, :] - self.means[None].astype(np.int64)) ** 2).sum(-1)       # find_distance (cluster.c:176-187)
        self.ids = d.argmin(1).astype(np.uint8)                                       # first minimum = lowest id on ties
        sums = _view(ptr, K * C + K, np.int64)
        for k in range(K):
            sums[k * C:(k + 1) * C] = x[self.ids == k].sum(0)
            sums[K * C + k] = int((self.ids == k).sum())

    def kmeans_update_dev(self, ptr):
        K, C = self.K, self.columns
        sums = _view(ptr, K * C + K, np.int64)
        counts = sums[K * C:].copy()
        if (counts == 0).any():
            raise RuntimeError("empty cluster")
        new = (sums[:K * C].reshape(K, C) // counts[:, None]).astype(np.uint8)        # recalculate_means (cluster.c:106-117)
        moved = ((new.astype(np.int64) - self.means.astype(np.int64)) ** 2).sum(1).astype(np.float64)
        self.means = new
        return moved, counts.astype(np.uint32)

    # the device-side loop decision of the library (update kernel), restated on the host
    def kmeans_update_async(self, ptr, threshold, max_iter):
        if not self.done:
            moved, counts = self.kmeans_update_dev(ptr)
            self.log.append(moved)
            self.last_counts = counts
            self.iters += 1
            if not (moved.max() > threshold) or self.iters >= max_iter:        # src/cluster.c:221-234
                self.done = True
        self.states.append((self.done, self.iters))

    def kmeans_poll(self, idx):
        return self.states[idx]

    def kmeans_result(self):
        return self.iters, np.array(self.log), self.last_counts

    def timings(self):
        return {"kmeans_assign_ms": 0.0, "kmeans_ms": 0.0}

    def kmeans_end(self, want_ids=True):
        return (self.ids if want_ids else None), self.means

    def cond_counts_len(self):
        return self.K * (1 + 72 * (self.columns - 1)) * 72

    def cond_counts_dev(self, ptr):
        out = _view(ptr, self.cond_counts_len(), np.int32)
        out[:] = self.O.cond_counts(self.rows, self.columns, self.K, self.ids).reshape(-1).view(np.int32)

    def quantize(self, tables, seed, **kw):
        return self.O.quantize(self.rows, self.columns, self.ids, tables, seed, first_line=self.first_line)

    def close(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, c, K, thr, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.bindings import DEBUG_SEED
        from tests.helpers import synthetic_tables
        rows = synth_rows(n, c, seed=77).numpy()
        b = shard_bounds(n, world)
        lo, hi = b[rank], b[rank + 1]
        local = np.ascontiguousarray(rows[lo:hi])
        fe = ShardedFrontEnd(handle=StandInHandle(), device="cpu")
        assert fe.world == world and fe.rank == rank
        fe.load_rows(local, hi - lo, c, c + 1, first_line=lo)
        picks = [(i * 7919 + 13) % n for i in range(K)]
        init = fe.broadcast_init_means(picks, local, lo, c)
        km = fe.kmeans(init, thr)
        counts = fe.cond_counts()
        q = fe.quantize(synthetic_tables(K, c, seed=5), DEBUG_SEED)
        np.savez(os.path.join(outdir, f"r{rank}.npz"), init=init, iters=km["iters"], ids=km["ids"], means=km["means"],
                 kcounts=km["counts"], moved=km["moved"], cond=counts, symbols=q["symbols"], line_err=q["line_err"],
                 allreduces=fe.allreduce_calls, lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,c,K,thr", [(2, 4001, 23, 3, 4.0), (3, 1503, 10, 2, 0.0)])
def test_sharded_equals_whole_on_gloo(tmp_path, oracle, world, n, c, K, thr):
    from oracle.bindings import DEBUG_SEED
    from tests.helpers import synthetic_tables
    mp.spawn(_worker, args=(world, _free_port(), n, c, K, thr, str(tmp_path)), nprocs=world, join=True)
    rows = synth_rows(n, c, seed=77).numpy()
    picks = [(i * 7919 + 13) % n for i in range(K)]
    whole = oracle.kmeans(rows, c, rows[picks, :c], thr)
    cond = oracle.cond_counts(rows, c, K, whole["ids"])
    wq = oracle.quantize(rows, c, whole["ids"], synthetic_tables(K, c, seed=5), DEBUG_SEED)
    parts = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    for p in parts:
        assert np.array_equal(p["init"], rows[picks, :c])
        assert int(p["iters"]) == whole["iters"]
        assert np.array_equal(p["means"], whole["means"])
        assert np.array_equal(p["kcounts"], whole["counts"])
        assert np.array_equal(p["moved"], whole["moved"])
        assert np.array_equal(p["cond"], cond)                                  # every rank holds the global table
        # one per iteration + one enqueued speculatively before the last outcome was known + the count tables
        assert int(p["allreduces"]) == whole["iters"] + 2
        assert int(p["lo"]) % 4 == 0
    assert np.array_equal(np.concatenate([p["ids"] for p in parts]), whole["ids"])
    assert np.array_equal(np.concatenate([p["symbols"] for p in parts]), wq["symbols"])
    assert np.array_equal(np.concatenate([p["line_err"] for p in parts]), wq["line_err"])


def test_shard_bounds():
    for n in (0, 1, 3, 4, 5, 1000, 1001, 20_000_000, 200_000_000):
        for w in (1, 2, 3, 4, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and len(b) == w + 1
            assert all(b[i] <= b[i + 1] for i in range(w))
            assert all(x % 4 == 0 or x == n for x in b[:-1])      # x == n: an empty trailing shard (tiny files)
    assert shard_bounds(200_000_000, 8) == [25_000_000 * r for r in range(9)]


def test_kmeans_pick_lines_matches_reference_formula():
    # glibc rand() seed-1 stream -> the lines the reference picks
    rand = [1804289383, 846930886, 1681692777, 1714636915, 1957747793, 424238335, 719885386, 1649760492,
            596516649, 1189641421]
    assert kmeans_pick_lines(1_000_000, 1, rand) == [930886]
    assert kmeans_pick_lines(20_000_000, 1, rand) == [3930886]
    assert kmeans_pick_lines(50_000_000, 3, rand) == [33930886, 27636915, 43238335]
    assert kmeans_pick_lines(200_000_000, 5, rand) == [183930886, 177636915, 193238335, 186760492, 49641421]
    assert kmeans_pick_lines(40_000_000, 2, rand) == [23930886, 17636915]
