"""Turn the raw ncu outputs under gpurun_out/ into the tracked summaries under profiles/ (round 2).

  python tools/make_profiles_r2.py r02 cfg4 "<profiled command>" <symbols per launch> gpurun_out/q_launches.csv gpurun_out/q_assign.ncu-rep ...

Writes profiles/<tag>_launches_<cfg>.csv (per-kernel device-time shares of one whole bench command) and
profiles/<tag>_ncu_full_<cfg>.json (the counters the roofline discussion uses, one record per captured launch, plus
dram bytes per symbol per kernel: what bench.py scales into roofline.traffic)."""
import collections
import csv
import json
import subprocess
import sys

tag, cfg, cmd, symbols, launches = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5]
reps = sys.argv[6:]

rows = [r for r in csv.reader(open(launches, errors="replace")) if len(r) > 5]
hdr = rows[0]
i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[i_val].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[i_unit], 1.0)
    a = agg.setdefault(r[i_name].split("(")[0].replace("void ", ""), [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(f"profiles/{tag}_launches_{cfg}.csv", "w") as f:
    f.write(f"# ncu launch list of `{cmd}`, kernels matching regex:qvz_\n")
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: compare SHARES\n")
    f.write(f"# total captured device time {tot/1e6:.2f} ms over {sum(a[0] for a in agg.values())} launches (warm-up step, timed step and the e2e step)\n")
    f.write("kernel,launches,total_us,share_pct,avg_us\n")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{n},{a[0]},{a[1]/1e3:.1f},{100*a[1]/tot:.2f},{a[1]/a[0]/1e3:.1f}\n")
    step = {n: a for n, a in agg.items() if any(t in n for t in ("kmeans_assign", "kmeans_update", "cond_counts_kernel", "cond_counts_planes", "quantize_batched", "draws_kernel", "f2_apply"))}
    st = sum(a[1] for a in step.values())
    f.write("# the kernels of the resident step alone (what `value` times): shares to compare with bench.py's stage_ms\n")
    f.write("kernel,share_of_step_pct\n")
    for n, a in sorted(step.items(), key=lambda x: -x[1][1]):
        f.write(f"{n},{100*a[1]/st:.2f}\n")

want = {"gpu__time_duration.sum": "time", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
        "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
        "smsp__inst_executed.sum": "warp_instructions", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier"}
out, per_symbol = [], {}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    idx = {k: i for i, k in enumerate(h)}

    def num(r, k):
        v, u = float(r[idx[k]].replace(",", "")), units[idx[k]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    for r in rr[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        d = {"kernel": name, "report": rep.split("/")[-1]}
        for k, short in want.items():
            if k in idx:
                d[f"{short} [{units[idx[k]]}]"] = r[idx[k]]
        d["warp_instructions_per_symbol_x32"] = round(32 * float(r[idx["smsp__inst_executed.sum"]].replace(",", "")) / symbols, 3)
        if num(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1.0) < 20.0:
            continue                                  # a speculative k-means launch after convergence: returns at once
        out.append(d)
        per_symbol.setdefault(name, []).append((num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")) / symbols)
json.dump({"command": cmd, "symbols_per_launch": symbols,
           "note": "ncu --set full --clock-control none --import-source on; one record per captured launch; "
                   "dram_bytes_per_symbol = (dram__bytes_read.sum + dram__bytes_write.sum) / symbols, averaged over the captured launches of a kernel",
           "launches": out, "dram_bytes_per_symbol": {k: sum(v) / len(v) for k, v in per_symbol.items()}},
          open(f"profiles/{tag}_ncu_full_{cfg}.json", "w"), indent=1)
print(open(f"profiles/{tag}_launches_{cfg}.csv").read())
print(json.dumps({k: sum(v) / len(v) for k, v in per_symbol.items()}, indent=1))
