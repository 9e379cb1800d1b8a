"""Summarise an .ncu-rep (one row per captured launch) into a small CSV for profiles/: duration, DRAM bytes, L2 hit rate,
issue-slot utilisation, occupancy, registers, grid, and the top stall reasons.   python tools/ncu_summary.py in.ncu-rep out.csv"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "lts__t_sectors.sum",
        "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch {i}" for i in range(len(rows) - 2)])
    for k in KEYS:
        if k in idx:
            w.writerow([k, units[idx[k]]] + [r[idx[k]] for r in rows[2:]])
    for k in stalls:
        vals = [r[idx[k]] for r in rows[2:]]
        try:
            if max(float(v) for v in vals) < 0.5:
                continue
        except ValueError:
            continue
        w.writerow([k.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", ""), "warps per issue"] + vals)
