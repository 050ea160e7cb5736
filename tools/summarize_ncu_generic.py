#!/usr/bin/env python
"""Any .ncu-rep -> a markdown table of the metrics the design discussion uses, one column per captured launch.
    python tools/summarize_ncu_generic.py gpurun_out/prof_decode.ncu-rep profiles/r01f_ncu_decode.md "title line"
"""
import csv
import io
import re
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]
ni = hdr.index("Kernel Name")


def short(name):
    m = re.search(r"(\w+_kernel<[^>]*>)", name)
    return "gb::" + m.group(1) if m else re.sub(r"\(.*", "", name)[-60:]


with open(out, "w") as f:
    f.write(f"# {title}\n\n")
    kn = [short(r[ni]) for r in data]
    f.write("| metric | unit | " + " | ".join(f"`{k}`" for k in kn) + " |\n|---|---|" + "---|" * len(kn) + "\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write(f"| {w} | {units[i]} | " + " | ".join(r[i] for r in data) + " |\n")
print("wrote", out)
