#!/usr/bin/env python
"""summarize_ncu_all.py REPORT.ncu-rep OUT.md [title] -- one table row per captured launch: duration, DRAM bytes and
throughput, occupancy, issue and pipe utilisation, the top stall reasons.  (Every second launch of
tools/profile_all_target.py: the first of each pair is the cold one.)"""
import csv
import io
import subprocess
import sys

rep, out_md = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]


def num(d, k):
    try:
        return float(d.get(k, "").replace(",", ""))
    except ValueError:
        return float("nan")


lines = [f"# {title}", "",
         "`ncu --set full --clock-control none --import-source on`; per launch (serialised, cold L2 -- shares and counters, not bench values).",
         "DRAM GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration; issue = smsp__issue_active; ALU / FMA = the two integer pipes",
         "(sm__inst_executed_pipe_alu / _fma, % of peak); stalls = warps stalled per issued instruction, top three.", "",
         "| kernel | grid x block | regs | us | DRAM MB (rd + wr) | DRAM GB/s | DRAM % of peak | warps/SMSP | issue % | ALU % | FMA % | waves/SM | stalls |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
seen = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "")
    key = (name, d.get("launch__grid_size"), d.get("launch__block_size"))
    seen[key] = seen.get(key, 0) + 1
    if seen[key] % 2 == 1 and "--all" not in sys.argv:
        continue   # first (cold) launch of the pair
    us = num(d, "gpu__time_duration.sum")
    rd, wr = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
    st = []
    for k, v in d.items():
        if "issue_stalled" in k and k.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    st = ", ".join(f"{n} {v:.1f}" for v, n in sorted(st, reverse=True)[:3])
    short = name.replace("void ", "").replace("gb::", "")
    lines.append(f"| `{short}` | {d.get('launch__grid_size')} x {d.get('launch__block_size')} | {d.get('launch__registers_per_thread')} | {us:.1f} | "
                 f"{rd:.1f} + {wr:.1f} | {(rd + wr) / us * 1e3:.0f} | {num(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{num(d, 'smsp__warps_active.avg.per_cycle_active'):.1f} | {num(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                 f"{num(d, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):.1f} | {num(d, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.1f} | "
                 f"{d.get('launch__waves_per_multiprocessor')} | {st} |")
open(out_md, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
