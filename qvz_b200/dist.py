"""Read-sharded front end: one process per GPU, contiguous line shards, torch.distributed for the exchange.

The three hot stages are independent per line; only integer sums cross ranks:

  k-means      per iteration one all-reduce(SUM) of int64[K*C + K] (column sums, then line counts); every
               rank then recentres redundantly and evaluates `moved > threshold` on identical integers, so
               all ranks leave the loop of do_kmeans_clustering (reference src/cluster.c:221-234) together.
  cond counts  one all-reduce(SUM) of the K*(1+72(C-1))*72 counters (reference src/codebook.c:193-205).
               They are uint32 in the reference (src/pmf.c:211-214); NCCL/gloo sum them as int32 -- the
               same bits in two's complement, and every true counter is <= N < 2^32.
  quantize     no exchange: the shard's first_line (a multiple of 4 => WELL word aligned) positions it in
               the reference's draw stream, draw = line*C + column (src/codebook.c:162-171).

Integer sums are order independent, so results are bit-identical for any world size.
There is no CPU path here: the stage calls go to the CUDA library handle (qvz_b200.lib.Handle).  The
`handle`/`device` arguments exist so that the host-side logic can be driven by the world_size-2 gloo
tests with a stand-in handle; the product always uses the defaults.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

ALPHABET = 72


def shard_bounds(n_lines: int, world: int) -> list[int]:
    """world+1 line boundaries of contiguous shards; inner boundaries are multiples of 4 lines so that
    every shard's first draw index first_line*C is a WELL word boundary for any C."""
    if world < 1:
        raise ValueError("world must be >= 1")
    b = [min(n_lines, ((n_lines * r // world) + 3) & ~3) for r in range(world)] + [n_lines]
    b[0] = 0
    for r in range(1, world + 1):
        b[r] = max(b[r], b[r - 1])
    return b


def kmeans_pick_lines(n_lines: int, K: int, rand_stream) -> list[int]:
    """Global line indices chosen by initialize_kmeans_clustering (reference src/cluster.c:199-201):
    block = rand() % block_count, line = rand() % blocks[block].count, blocks of 1 000 000 lines
    (MAX_LINES_PER_BLOCK, include/lines.h:12), the last block holding the remainder (src/lines.c:88-126)."""
    per = 1_000_000
    blocks = (n_lines + per - 1) // per
    out = []
    it = iter(rand_stream)
    for _ in range(K):
        b = next(it) % blocks
        cnt = per if b + 1 < blocks or n_lines % per == 0 else n_lines % per
        out.append(b * per + next(it) % cnt)
    return out


class ShardedFrontEnd:
    """k-means / conditional counts / quantize over one shard per rank."""

    def __init__(self, local_device: int = 0, handle=None, device=None, group=None):
        if handle is None:
            from . import lib
            handle = lib.Handle(local_device)          # raises without a GPU / without the built library
            device = torch.device("cuda", local_device)
        self.h = handle
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._sums = None
        self._counts = None
        self.allreduce_calls = 0
        self._ext = None                                       # the library's stream as a torch stream (built once)
        self.kmeans_ms = self.kmeans_assign_ms = 0.0           # device time of the last kmeans() (CUDA events, lib stream)

    # the library works on its own stream: run the collective on that stream so no host sync is needed
    def _lib_stream(self):
        if self._ext is None:
            self._ext = torch.cuda.ExternalStream(self.h.stream, device=self.device)
        return self._ext

    def _on_lib_stream(self):
        if self.device.type == "cuda":
            return torch.cuda.stream(self._lib_stream())
        import contextlib
        return contextlib.nullcontext()

    def _allreduce(self, t: torch.Tensor):
        if self.world > 1:
            with self._on_lib_stream():
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.allreduce_calls += 1

    def load_rows(self, rows, n_lines: int, columns: int, row_stride: int, first_line: int):
        if first_line & 3:
            raise ValueError("a shard must start on a multiple of 4 lines (WELL word boundary)")
        self.h.load_rows(rows, n_lines, columns, row_stride, first_line=first_line)

    def broadcast_init_means(self, picks, local_rows: np.ndarray, first_line: int, columns: int) -> np.ndarray:
        """K x C initial centroids = the picked global lines; each is broadcast by the rank that holds it."""
        n_local = local_rows.shape[0]
        init = np.zeros((len(picks), columns), np.uint8)
        for j, gl in enumerate(picks):
            mine = first_line <= gl < first_line + n_local
            buf = torch.zeros(columns, dtype=torch.uint8, device=self.device)
            if mine:
                buf.copy_(torch.from_numpy(np.ascontiguousarray(local_rows[gl - first_line, :columns])))
            if self.world > 1:
                # exactly one rank holds the line: a SUM of one row and zeros is that row
                w = buf.to(torch.int32)
                dist.all_reduce(w, op=dist.ReduceOp.SUM, group=self.group)
                buf = w.to(torch.uint8)
            init[j] = buf.cpu().numpy()
        return init

    def kmeans(self, init_means: np.ndarray, threshold: float = 4.0, max_iter: int = 1000, want_ids: bool = True):
        """do_kmeans_clustering (reference src/cluster.c:212-244) over all shards."""
        init_means = np.ascontiguousarray(init_means, dtype=np.uint8)
        K, C = init_means.shape
        h = self.h
        h.kmeans_begin(init_means)
        if self._sums is None or self._sums.numel() != K * C + K:
            # torch.empty, not zeros: the library zeroes / overwrites the buffer itself on ITS stream, and a fill queued
            # on torch's current stream would not be ordered against that
            self._sums = torch.empty(K * C + K, dtype=torch.int64, device=self.device)
        sums = self._sums
        cuda = self.device.type == "cuda"

        def enqueue():
            h.kmeans_assign_dev(sums.data_ptr())               # local assignment + local integer sums
            self._allreduce(sums)                              # global sums (exact), on the library's stream
            h.kmeans_update_async(sums.data_ptr(), threshold, max_iter)    # recalculate_means + loop decision, on the device

        # Iteration i+1 is enqueued before anybody looks at the outcome of iteration i.  Every rank sees the same sums,
        # hence the same decision: all of them enqueue the same collectives in the same order, and launches made after
        # the run has ended return at once (src/cluster.c:221-234 evaluated in the update kernel).
        i = 0
        if max_iter > 0:
            enqueue()
            while True:
                if i + 1 < max_iter:
                    enqueue()
                done, _ = h.kmeans_poll(i)
                if done or i + 1 >= max_iter:
                    break
                i += 1
        iters, moved_log, counts = h.kmeans_result()
        if cuda:
            tm = h.timings()
            self.kmeans_assign_ms, self.kmeans_ms = tm["kmeans_assign_ms"], tm["kmeans_ms"]
        ids, means = h.kmeans_end(want_ids=want_ids)
        return dict(iters=iters, ids=ids, means=means, counts=counts, moved=np.asarray(moved_log))

    def cond_counts(self, want_host: bool = True):
        """calculate_statistics' counting loop (reference src/codebook.c:193-205) over all shards;
        returns the global table [K, 1+72(C-1), 72] on every rank (or None)."""
        h = self.h
        n = h.cond_counts_len()
        if self._counts is None or self._counts.numel() != n:
            self._counts = torch.empty(n, dtype=torch.int32, device=self.device)     # zeroed by the library on its stream
        h.cond_counts_dev(self._counts.data_ptr())
        self._allreduce(self._counts)
        if not want_host:
            return None
        if self.device.type == "cuda":
            with self._on_lib_stream():
                host = self._counts.cpu()
        else:
            host = self._counts.clone()
        return host.numpy().view(np.uint32).reshape(h.K, 1 + ALPHABET * (h.columns - 1), ALPHABET)

    def quantize(self, tables, seed, **kw):
        """The shard's part of start_qv_compression's walk; outputs are this rank's lines, in line order."""
        return self.h.quantize(tables, seed, **kw)

    def close(self):
        self.h.close()
