#!/usr/bin/env python
"""bench.py -- Gsymbols/s of the qvz front end (k-means + conditional counts + quantize walk) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--lines L]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic
quality lines:
  value      whole-job Gsymbols/s with the rows already resident in HBM (device time, CUDA events on the
             library's stream, max over ranks);
  e2e        the same metric through the C ABI with HOST buffers: H2D of the rows, the three stage calls
             with their D2H results (cluster ids, count tables, symbol stream) inside the timed region;
  roofline   algorithmic bytes / CUDA-event duration of the dominant kernel vs the measured HBM peak;
  cpu_baseline  the reference algorithm on the host cores on a bounded sample (rank 0, N=1 only).
--impl reference times the reference's own CPU implementation (oracle/_ref, else the oracle port).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gsymbols/s cluster+PMF+quantize"
UNIT = "Gsymbols/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------ workload
def workload(args, world):
    """cfg2 of BASELINE.json at N=1 (the largest single-GPU configuration that the metric is quoted on);
    under torchrun every rank gets the same per-GPU shard size (weak scaling)."""
    from qvz_b200.synth import CONFIGS
    cfg = dict(CONFIGS[args.config])
    if args.lines:
        cfg["lines"] = args.lines
    cfg["name"] = args.config
    return cfg


DIST = {"M": 2, "L": 3, "A": 1}        # include/distortion.h:7-9
# glibc rand() after the implicit srand(1): what the reference's k-means initialisation consumes
GLIBC_RAND_SEED1 = [1804289383, 846930886, 1681692777, 1714636915, 1957747793, 424238335, 719885386, 1649760492,
                    596516649, 1189641421]


def make_tables(cfg, counts):
    """Quantizer tables for the quantize stage = what the reference's encode() would hand to start_qv_compression:
    the host codebook designer (qvz_b200/host, bit-exact with generate_codebooks) run on the conditional counts of
    this very workload, with the config's -f/-r target and -d distortion."""
    from qvz_b200 import hostlib
    mode = hostlib.MODE_RATIO if cfg["mode"] == "ratio" else hostlib.MODE_FIXED
    t0 = time.time()
    cb = hostlib.design_codebooks(counts, cfg["columns"], cfg["clusters"], mode, cfg["ratio"], DIST[cfg["dist"]])
    log(f"host codebook design took {time.time()-t0:.1f}s (outside the metric, like generate_codebooks)")
    return cb, "lloyd-max: host codebook design (bit-exact with generate_codebooks) on this workload's GPU counts"


class ClockSampler:
    """nvidia-smi SM clock / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def bind_to_gpu_numa_node(gpu_index):
    """Run this rank on the CPUs closest to its GPU (NVML's CPU affinity), so that its pinned host buffers are
    first-touched on that NUMA node and the H2D/D2H copies of the e2e leg do not cross sockets.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(gpu_index), (os.cpu_count() + 63) // 64)
        near = {i for i in range(os.cpu_count()) if (mask[i // 64] >> (i % 64)) & 1} & os.sched_getaffinity(0)
        if near:
            os.sched_setaffinity(0, near)
            return len(near)
    except Exception as e:                         # no NVML, restricted cgroup, ...: keep the default placement
        log(f"[gpu {gpu_index}] NUMA binding skipped: {e!r}")
    return 0


# ------------------------------------------------------------------------------------------ native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    from qvz_b200 import lib
    from qvz_b200.synth import synth_rows

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        bind_to_gpu_numa_node(local)               # before any pinned allocation: first touch decides the NUMA node
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = workload(args, world)
    n, c, k = cfg["lines"], cfg["columns"], cfg["clusters"]
    first_line = rank * n                          # shard `rank` of a world*n-line file (n % 4 == 0)
    sym_per_rank = n * c

    t0 = time.time()
    rows_d = synth_rows(n, c, seed=1234 + rank, profile=cfg["profile"], device="cuda")
    rows = torch.empty(rows_d.shape, dtype=torch.uint8, pin_memory=True)
    rows.copy_(rows_d)
    torch.cuda.synchronize()
    del rows_d
    torch.cuda.empty_cache()
    rows_np = rows.numpy()
    log(f"[rank {rank}] synthetic {n}x{c} ready in {time.time()-t0:.1f}s")

    if world > 1:
        from qvz_b200.dist import ShardedFrontEnd
        fe = ShardedFrontEnd(local)
    else:
        fe = None
    h = fe.h if fe else lib.Handle(local)

    # initial centroids: the rows picked by initialize_kmeans_clustering's unseeded rand() stream (src/cluster.c:199-201)
    from qvz_b200.dist import kmeans_pick_lines
    total = n * world
    picks = kmeans_pick_lines(total, k, GLIBC_RAND_SEED1)
    if fe:
        init = fe.broadcast_init_means(picks, rows_np, first_line, c)
    else:
        init = np.ascontiguousarray(rows_np[picks, :c])
    seed = np.full(32, 0x55555555, np.uint32)      # the reference's DEBUG seed (src/qv_stream.c:82)

    # resident inputs + tables (outside the timed region)
    h.load_rows(rows, n, c, c + 1, first_line=first_line)
    if fe:
        km = fe.kmeans(init, cfg.get("threshold", 4.0), want_ids=False)
        counts = fe.cond_counts(want_host=True)
    else:
        km = h.kmeans(init, cfg.get("threshold", 4.0), want_ids=False)
        counts = h.cond_counts()
    tables, tables_kind = make_tables(cfg, counts)
    tstruct = tables.tables                         # struct qvz_flat_tables view of the designed codebooks
    log(f"[rank {rank}] k-means iterations {km['iters']}, tables: {tables_kind}")

    ids_host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    sym_host = torch.empty((n, c), dtype=torch.uint8, pin_memory=True)
    stream = torch.cuda.ExternalStream(h.stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident():
        if args.prefetch:
            h.prefetch_draws(seed)                 # the draws depend on the seed only: generated under k-means
        if fe:
            fe.kmeans(init, cfg.get("threshold", 4.0), want_ids=False)
            fe.cond_counts(want_host=False)
        else:
            h.kmeans(init, cfg.get("threshold", 4.0), want_ids=False)
            h.cond_counts(want=False)
        h.quantize(tstruct, seed, want_symbols=False)
        tm = h.timings()
        if fe:                                     # the stepping calls are timed by the sharded front end
            tm["kmeans_ms"], tm["kmeans_assign_ms"], tm["kmeans_iters"] = fe.kmeans_ms, fe.kmeans_assign_ms, km["iters"]
        return tm

    def step_e2e():
        h.load_rows(rows, n, c, c + 1, first_line=first_line)
        if fe:
            fe.kmeans(init, cfg.get("threshold", 4.0), want_ids=True)
            fe.cond_counts(want_host=True)
        else:
            h.kmeans(init, cfg.get("threshold", 4.0), ids_out=ids_host.numpy())
            h.cond_counts()
        h.quantize(tstruct, seed, symbols_out=sym_host)

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    h.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = {"kmeans_ms": 0.0, "kmeans_assign_ms": 0.0, "cond_counts_ms": 0.0, "quantize_ms": 0.0, "quantize_draws_ms": 0.0,
             "quantize_setup_ms": 0.0}
    ev0.record(stream)
    for _ in range(args.steps):
        tm = step_resident()
        for key in stage:
            stage[key] += tm[key]
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = h.timings()["kernel_launches"]
    iters = tm["kmeans_iters"]
    clocks = sampler.stop() if rank == 0 else None

    # e2e: host buffers in, host buffers out
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0

    # Extra information, not the headline: the same calls with TWO jobs in flight on this GPU (two handles, two host
    # threads; every job still copies its rows in and its symbols out).  One job cannot overlap its own H2D and D2H --
    # the symbols exist only after the whole file has been counted -- but PCIe is full duplex, so job A's upload hides
    # behind job B's download.  This is what compressing a list of files looks like.  (Measured on this pool: 29.4 vs
    # 26.5 Gsymbols/s -- the two directions overlap far less than full duplex would allow; opt-in, --two-jobs.)
    two_jobs = None
    if world == 1 and args.two_jobs:
        import threading
        h2 = lib.Handle(local)
        ids2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        sym2 = torch.empty((n, c), dtype=torch.uint8, pin_memory=True)

        upload = threading.Lock()                  # one upload at a time: the jobs fall into upload/download alternation

        def job(hh, ids_buf, sym_buf, reps):
            for _ in range(reps):
                with upload:
                    hh.load_rows(rows, n, c, c + 1, first_line=first_line)
                hh.kmeans(init, cfg.get("threshold", 4.0), ids_out=ids_buf.numpy())
                hh.cond_counts()
                hh.quantize(tstruct, seed, symbols_out=sym_buf)

        job(h2, ids2, sym2, 1)                     # warm the second handle (buffers, jump tables)
        torch.cuda.synchronize()
        th = [threading.Thread(target=job, args=(h, ids_host, sym_host, args.steps)),
              threading.Thread(target=job, args=(h2, ids2, sym2, args.steps))]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        same = bool(torch.equal(sym_host, sym2))
        two_jobs = {"value": round(2 * args.steps * sym_per_rank / dt / 1e9, 4), "unit": UNIT, "jobs_in_flight": 2,
                    "ms_per_job": round(dt / (2 * args.steps) * 1e3, 2), "outputs_identical": same}
        h2.close()
    h2d = n * (c + 1) + int(np.asarray(tables.qmap).size * 4)
    d2h = n * c + n + (counts.nbytes if counts is not None else 0)

    if world > 1:
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    for key in stage:
        stage[key] /= args.steps
    total_sym = sym_per_rank * world
    value = total_sym / (ms_step * 1e-3) / 1e9
    e2e_value = total_sym / (e2e_s / args.steps) / 1e9

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # per-kernel durations (CUDA events on the library's streams, averaged over the timed steps) and algorithmic bytes
    # (1 B per QV read, 1 B per cluster id, 1 B per emitted symbol; the WELL draws are an
    # intermediate, not algorithmic I/O -- the draw generator's time is shown but it has no roofline of its own)
    walk_ms = stage["quantize_ms"] - stage["quantize_draws_ms"]
    if k == 1:
        # one cluster: the k-means stage IS the counting pass (its column sums are marginals of the count table, csrc/kmeans.cu);
        # qvz_gpu_cond_counts then hands out the table that pass left on the device
        counts_ms = stage["kmeans_assign_ms"] + stage["cond_counts_ms"]
        kern = {"cond_counts": (counts_ms, sym_per_rank + n), "quantize_walk": (walk_ms, 2 * sym_per_rank + n)}
        share = {"cond_counts": counts_ms, "quantize_walk": walk_ms}
    else:
        km_launches = max(iters, 1)
        kern = {"kmeans_assign": (stage["kmeans_assign_ms"] / km_launches, sym_per_rank + n),
                "cond_counts": (stage["cond_counts_ms"], sym_per_rank + n),
                "quantize_walk": (walk_ms, 2 * sym_per_rank + n)}
        share = {"kmeans_assign": stage["kmeans_assign_ms"], "cond_counts": stage["cond_counts_ms"], "quantize_walk": walk_ms}
    dom = max(share, key=share.get)                # dominant kernel = largest share of the step
    dur_ms, alg_bytes = kern[dom]
    achieved = alg_bytes / (dur_ms * 1e-3) / 1e9
    # one cluster and Q <= 41 (this workload): the counting pass is the lane-private byte-plane kernel
    names = {"kmeans_assign": "qvz_kmeans_assign_kernel", "cond_counts": "qvz_cond_counts_planes_kernel" if k == 1 else "qvz_cond_counts_kernel",
             "quantize_walk": "qvz_quantize_batched_kernel"}
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": ncu_traffic(dom, cfg),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "per_kernel_GBps": {names[kk]: round(b / (d * 1e-3) / 1e9, 1) for kk, (d, b) in kern.items() if d > 0},
                "per_kernel_ms": {**{names[kk]: round(d, 4) for kk, (d, b) in kern.items()}, "qvz_draws_kernel": round(stage["quantize_draws_ms"], 4)},
                "quantize_stage_GBps_incl_draw_generator": round((2 * sym_per_rank + n) / (stage["quantize_ms"] * 1e-3) / 1e9, 1),
                "algorithmic_bytes_per_launch": alg_bytes,
                # the whole step by SURVEY.md section 8d's accounting: (I+3)*N*C + (I+2)*N algorithmic bytes for I k-means iterations
                "whole_step": {"algorithmic_bytes": (iters + 3) * sym_per_rank + (iters + 2) * n,
                               "achieved": round(((iters + 3) * sym_per_rank + (iters + 2) * n) / (ms_step * 1e-3) / 1e9, 1),
                               "frac": round(((iters + 3) * sym_per_rank + (iters + 2) * n) / (ms_step * 1e-3) / 1e9 / peak, 4)}}

    cpu = cpu_baseline(cfg, args) if world == 1 and not args.no_cpu else None

    out = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": {"workload": f"{cfg['name']}: {n} lines x {c} columns per GPU, {k} cluster(s), "
                                  f"{'-f' if cfg['mode']=='ratio' else '-r'} {cfg['ratio']} -d {cfg['dist']}",
                      "lines_per_gpu": n, "columns": c, "clusters": k, "kmeans_iterations": iters,
                      "tables": tables_kind, "l2": "inputs (%.2f GB per GPU) exceed the 126 MB L2" % (n * c / 1e9),
                      "sharding": "contiguous line shards, NCCL all-reduce of int64 centroid sums and uint32 counts" if world > 1 else "single GPU"},
           "stage_ms": {kk: round(v, 4) for kk, v in stage.items()},
           "e2e": {"value": round(e2e_value, 4), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": round(e2e_s / args.steps * 1e3, 2),
                   **({"two_jobs_in_flight": two_jobs} if two_jobs else {})},
           "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks}
    if cpu:
        out["cpu_baseline"] = cpu
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(kernel, cfg):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture of this very workload (profiles/*_ncu_full_cfg2.json); None when the workload is not the captured one."""
    import glob
    if (cfg["name"], cfg["lines"], cfg["columns"], cfg["clusters"]) != ("cfg2", 20_000_000, 150, 1):
        return None
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_full_cfg2.json")))
    if not files:
        return None
    per = json.load(open(files[-1]))["dram_bytes_per_launch"]
    pick = {"quantize_walk": ("qvz_quantize_batched_kernel",), "cond_counts": ("qvz_cond_counts_kernel", "qvz_cond_counts_planes_kernel"),
            "kmeans_assign": ("qvz_kmeans_assign_kernel",)}[kernel]
    tot = sum(v for k, v in per.items() if any(p in k for p in pick))
    return int(tot) if tot else None


# ------------------------------------------------------------------------------------------ CPU legs
def _oracle_tables(cb):
    """The designed codebooks as the checker's FlatTables (same arrays, no copy)."""
    from oracle.bindings import FlatTables
    return FlatTables(cb.clusters, cb.columns, cb.nctx, cb.ctx_of, cb.q_off, cb.qratio, cb.qmap, cb.smap, cb.distortion)


def _cpu_sample(cfg, lines):
    from qvz_b200.synth import synth_rows
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    return synth_rows(lines, cfg["columns"], seed=1234, profile=cfg["profile"], device=dev).cpu().numpy()


def cpu_baseline(cfg, args):
    """Reference algorithm on the host: the oracle port's three stages on a bounded sample."""
    from oracle.bindings import Oracle
    lines = min(cfg["lines"], args.cpu_lines)
    c, k = cfg["columns"], cfg["clusters"]
    rows = _cpu_sample(cfg, lines)
    O = Oracle()
    init = rows[[(i * 104_729 + 930_886) % lines for i in range(k)], :c]
    seed = np.full(32, 0x55555555, np.uint32)
    t0 = time.perf_counter()
    km = O.kmeans(rows, c, init, cfg.get("threshold", 4.0))
    t1 = time.perf_counter()
    counts = O.cond_counts(rows, c, k, km["ids"])
    t2 = time.perf_counter()
    tables, _ = make_tables(cfg, counts)          # codebook design: outside the metric on both arms
    t2b = time.perf_counter()
    O.quantize(rows, c, km["ids"], _oracle_tables(tables), seed, want_qv=False, want_err=True)
    t3 = time.perf_counter()
    t3 -= t2b - t2
    return {"value": round(lines * c / (t3 - t0) / 1e9, 5), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {lines} lines of the workload ({t3-t0:.1f} s: kmeans {t1-t0:.2f} counts {t2-t1:.2f} quantize {t3-t2:.2f})"}


def run_reference(args):
    """The reference's own single-threaded C on the host cores: unmodified do_kmeans loop functions,
    calculate_statistics and the choose_quantizer walk from oracle/_ref/libqvzref.so."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    from oracle.bindings import MODE_FIXED, MODE_RATIO, Oracle, Ref, ref_available
    cfg = workload(args, world)
    lines = min(cfg["lines"], args.cpu_lines)
    c, k = cfg["columns"], cfg["clusters"]
    rows = _cpu_sample(cfg, lines)
    picks = [(i * 104_729 + 930_886) % lines for i in range(k)]
    seed = np.full(32, 0x55555555, np.uint32)
    times = []
    if ref_available():
        R = Ref()
        kind = "reference"
        mode = MODE_RATIO if cfg["mode"] == "ratio" else MODE_FIXED
        mk = lambda: R.session(rows, c, k, threshold=cfg.get("threshold", 4.0), mode=mode, ratio=cfg["ratio"], distortion=DIST[cfg["dist"]])
        s0 = mk()
        s0.kmeans(picks)
        t0 = time.perf_counter()
        s0.tables()                                  # generate_codebooks: outside the metric (host codebook design)
        log(f"reference codebook design took {time.perf_counter()-t0:.1f}s (not part of the metric)")
        for i in range(args.warmup + args.steps):
            s = mk()
            t0 = time.perf_counter()
            s.kmeans(picks)
            s.stats()
            t1 = time.perf_counter()
            s0.quantize(seed, want_qv=False, want_err=True)
            t2 = time.perf_counter()
            if i >= args.warmup:
                times.append(t2 - t0)
    else:
        O = Oracle()
        kind = "port"
        init = rows[picks, :c]
        km = O.kmeans(rows, c, init, cfg.get("threshold", 4.0))
        tables, _ = make_tables(cfg, O.cond_counts(rows, c, k, km["ids"]))
        otab = _oracle_tables(tables)
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            km = O.kmeans(rows, c, init, cfg.get("threshold", 4.0))
            O.cond_counts(rows, c, k, km["ids"])
            O.quantize(rows, c, km["ids"], otab, seed, want_qv=False, want_err=True)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    value = lines * c / sec / 1e9
    sample = f"{lines} lines x {c} columns per step (bounded sample of {cfg['name']}), single thread: the reference has no threading"
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": {"workload": f"{cfg['name']}: {cfg['lines']} lines x {c} columns per GPU, {k} cluster(s), "
                                  f"{'-f' if cfg['mode']=='ratio' else '-r'} {cfg['ratio']} -d {cfg['dist']}",
                      "lines_per_gpu": cfg["lines"], "columns": c, "clusters": k,
                      "tables": "the reference's own generate_codebooks on the sample", "sharding": "single host thread"},
           "cpu_baseline": {"value": round(value, 5), "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
           "e2e": {"value": round(value, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--lines", type=int, default=0, help="override lines per GPU (parity/dev runs)")
    ap.add_argument("--cpu-lines", type=int, default=1_000_000, help="lines in the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--two-jobs", action="store_true", help="also measure e2e with two jobs in flight on the GPU (extra information)")
    ap.add_argument("--prefetch", type=int, default=0, help="1: start the WELL draw generation at the start of the step (overlaps k-means)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
