#!/usr/bin/env python
"""ncu_brief.py REPORT.ncu-rep [kernel-regex] -- the handful of counters that say why a kernel is below its roofline."""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    if pat and not pat.search(d.get("Kernel Name", "")):
        continue
    print("##", d.get("Kernel Name"), "id", d.get("ID"))
    for k in KEYS:
        if k in d:
            print(f"  {k:72s} {d[k]} {u.get(k, '')}")
    st = []
    for k, v in d.items():
        if "issue_stalled" in k and k.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("  stalls (warps per issue):", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
