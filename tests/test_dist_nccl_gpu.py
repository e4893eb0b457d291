"""The read-sharded front end over NCCL on real GPUs: one process per GPU, the library's kernels on each shard,
all-reduce of the integer sums on the library's stream (qvz_b200/dist.py).  The result of the distributed run
must equal the CPU checker on the whole file.  Needs >= 2 GPUs (skipped otherwise); the CPU version of this
test, with a stand-in handle over gloo, is tests/test_dist_gloo.py."""
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

from qvz_b200.dist import ShardedFrontEnd, shard_bounds  # noqa: E402
from qvz_b200.synth import synth_rows  # noqa: E402

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, c, K, thr, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle.bindings import DEBUG_SEED
        from tests.helpers import synthetic_tables
        rows = synth_rows(n, c, seed=77).numpy()
        b = shard_bounds(n, world)
        lo, hi = b[rank], b[rank + 1]
        local = np.ascontiguousarray(rows[lo:hi])
        fe = ShardedFrontEnd(rank)                               # the CUDA library on device `rank`
        fe.load_rows(local, hi - lo, c, c + 1, first_line=lo)
        picks = [(i * 7919 + 13) % n for i in range(K)]
        init = fe.broadcast_init_means(picks, local, lo, c)
        tables = synthetic_tables(K, c, seed=5)
        for rep in range(2):                                     # the second run reuses every persistent buffer
            km = fe.kmeans(init, thr)
            counts = fe.cond_counts()
            q = fe.quantize(tables, DEBUG_SEED, want_err=True)
        np.savez(os.path.join(outdir, f"r{rank}.npz"), init=init, iters=km["iters"], ids=km["ids"], means=km["means"],
                 kcounts=km["counts"], moved=km["moved"], cond=counts, symbols=q["symbols"], line_err=q["line_err"])
        fe.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,c,K,thr", [(2, 200_003, 101, 3, 4.0), (2, 50_001, 36, 1, 4.0)])
def test_sharded_equals_whole_on_nccl(tmp_path, oracle, world, n, c, K, thr):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
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
        assert np.array_equal(p["cond"], cond)
    assert np.array_equal(np.concatenate([p["ids"] for p in parts]), whole["ids"])
    assert np.array_equal(np.concatenate([p["symbols"] for p in parts]), wq["symbols"])
    assert np.array_equal(np.concatenate([p["line_err"] for p in parts]), wq["line_err"])
