"""Print a compact per-kernel summary from `ncu -i X.ncu-rep --page raw --csv` output (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr, units = rows[0], rows[1]
want = {'gpu__time_duration.sum': 'time', 'dram__bytes_read.sum': 'dram_rd', 'dram__bytes_write.sum': 'dram_wr',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram%', 'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm%',
        'l1tex__t_sector_hit_rate.pct': 'l1hit%', 'lts__t_sector_hit_rate.pct': 'l2hit%',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'lts%', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed': 'l1tex%',
        'sm__warps_active.avg.pct_of_peak_sustained_active': 'occ%', 'launch__registers_per_thread': 'regs',
        'smsp__inst_executed.sum': 'inst', 'launch__grid_size': 'grid', 'launch__block_size': 'block',
        'launch__occupancy_limit_shared_mem': 'occlim_smem', 'launch__occupancy_limit_registers': 'occlim_regs',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_conflicts',
        'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'st_long_sb',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'st_short_sb',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio': 'st_lg_thr',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio': 'st_mio_thr',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio': 'st_barrier',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio': 'st_wait',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio': 'st_math',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio': 'st_notsel',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio': 'st_dispatch',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio': 'st_branch',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio': 'st_noinst',
        'lts__t_bytes.sum': 'l2_bytes', 'l1tex__t_bytes.sum': 'l1_bytes', 'sm__cycles_active.avg': 'sm_cycles'}
idx = {h: i for i, h in enumerate(hdr)}
seen = {}
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0][-60:]
    seen[name] = seen.get(name, 0) + 1
    if seen[name] > int(sys.argv[2]) if len(sys.argv) > 2 else seen[name] > 1:
        continue
    print('==', name)
    out = []
    for k, short in want.items():
        if k in idx:
            v = r[idx[k]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            out.append(f"{short}={v}{units[idx[k]] if units[idx[k]] not in ('%', '') else ''}")
    print('  ' + '  '.join(out))
