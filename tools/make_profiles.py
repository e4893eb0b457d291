"""Turn the raw ncu outputs under gpurun_out/ into the tracked summaries under profiles/ (named per round).

  python tools/make_profiles.py r01 gpurun_out/launches_r1.csv gpurun_out/prof_r1c.ncu-rep "<command that was profiled>"
"""
import collections
import csv
import json
import subprocess
import sys

tag, launches, rep, cmd = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]

# ---- launch list: every launch of our kernels with its device time -> per-kernel totals and shares
rows = [r for r in csv.reader(open(launches, errors="replace")) if len(r) > 5]
hdr = rows[0]
i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[i_val].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(r[i_name].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(f"profiles/{tag}_launches_bench_cfg2.csv", "w") as f:
    f.write(f"# ncu launch list of `{cmd}` (cfg2: 20M x 150, K=1), kernels matching regex:qvz_\n")
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: compare SHARES\n")
    f.write(f"# total captured device time {tot/1e6:.2f} ms over {sum(a[0] for a in agg.values())} launches\n")
    f.write("kernel,launches,total_us,share_pct,avg_us\n")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{n},{a[0]},{a[1]/1e3:.1f},{100*a[1]/tot:.2f},{a[1]/a[0]/1e3:.1f}\n")

# ---- full capture: the metrics the roofline numbers come from, one line per captured launch
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
idx = {k: i for i, k in enumerate(h)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out, traffic = [], {}
for r in rr[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    d = {"kernel": name}
    for k in want:
        if k in idx:
            d[k + " [" + units[idx[k]] + "]"] = r[idx[k]]
    out.append(d)
    def num(k):
        v, u = float(r[idx[k]].replace(",", "")), units[idx[k]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    traffic[name] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
json.dump({"command": cmd, "note": "ncu --set full --clock-control none, cfg2 (20M x 150, K=1); dram bytes per launch", "launches": out,
           "dram_bytes_per_launch": traffic}, open(f"profiles/{tag}_ncu_full_cfg2.json", "w"), indent=1)
print(open(f"profiles/{tag}_launches_bench_cfg2.csv").read())
print(json.dumps(traffic, indent=1))
