"""Summarise an .ncu-rep (read here, no GPU needed) into a text file for profiles/.
usage: ncu_summary.py REPORT.ncu-rep OUT.txt [kernel-regex]"""
import csv, io, re, subprocess, sys, collections

rep, out = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__waves_per_multiprocessor",
        "sm__maximum_warps_per_active_cycle_pct"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
lines = []
for r in data:
    name = r[col["Kernel Name"]]
    if pat and not pat.search(name):
        continue
    lines.append("== launch id %s  %s  grid %s block %s" % (r[col["ID"]], name[:100], r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]))
    for k in KEYS:
        if k in col:
            lines.append("  %-70s %-12s %s" % (k, units[col[k]], r[col[k]]))
    stalls = []
    for h in hdr:
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h) or \
            re.match(r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", h)
        if m:
            try:
                stalls.append((float(r[col[h]]), m.group(1)))
            except ValueError:
                pass
    for v, nm in sorted(stalls, reverse=True)[:8]:
        lines.append("  stall %-30s %.3f" % (nm, v))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
