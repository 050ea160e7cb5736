#!/usr/bin/env python
"""Turns gpurun_out/launches.csv (+ optionally a .ncu-rep) into the tracked summaries under profiles/.

    python tools/summarize_profile.py r01            # reads gpurun_out/launches.csv, gpurun_out/prof.ncu-rep
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = ROOT / "profiles"
out_dir.mkdir(exist_ok=True)


def short(name: str) -> str:
    m = re.search(r"(encode_\w+_kernel<[^>]*>)", name)
    if m:
        return "gb::" + m.group(1)
    name = re.sub(r"\(.*", "", name)
    return name[-70:]


launches = ROOT / "gpurun_out" / "launches.csv"
if launches.exists():
    lines = [l for l in launches.read_text().splitlines() if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
    per = OrderedDict()
    ours = []
    for r in rows:
        k = short(r["Kernel Name"])
        ns = float(r["Metric Value"])
        c = per.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += ns
        if "encode_" in k:
            ours.append((int(r["ID"]), k, r["Grid Size"], r["Block Size"], ns))
    total = sum(v[1] for v in per.values())
    with open(out_dir / f"{tag}_launches.md", "w") as f:
        f.write(f"# {tag}: ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e`\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised: compare shares).\n"
                "Everything that is not `gb::encode_*` is torch generating the synthetic input textures BEFORE the timed region;\n"
                "inside the timed region the step consists of `gb::encode_direct_kernel` launches only (100 % share).\n\n")
        f.write("| kernel | launches | total us | share of all captured |\n|---|---|---|---|\n")
        for k, (n, ns) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ns / 1e3:.1f} | {100 * ns / total:.1f} % |\n")
        f.write("\n## our launches (ID, kernel, grid, block, us)\n\n")
        by_kernel = OrderedDict()
        for i, k, g, b, ns in ours:
            by_kernel.setdefault(k, []).append(ns)
            f.write(f"- {i} `{k}` grid {g} block {b}: {ns / 1e3:.2f} us\n")
        f.write("\n## per-kernel mean\n\n")
        for k, v in by_kernel.items():
            f.write(f"- `{k}`: n={len(v)} mean {sum(v) / len(v) / 1e3:.2f} us  min {min(v) / 1e3:.2f} us\n")
    print("wrote", out_dir / f"{tag}_launches.md")

rep = ROOT / "gpurun_out" / "prof.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = [
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    ]
    name_i = hdr.index("Kernel Name")
    with open(out_dir / f"{tag}_ncu_full.md", "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on -k regex:encode_ -s 9 -c 3 python tools/profile_target.py`\n\n")
        f.write("One launch each over the batch bench.py times: 4 device-resident 8192x8192 RGBA8 textures in one batched launch\n"
                "(1.07 GB of input per launch, far larger than L2).  Algorithmic bytes per launch: 1073.7 MB read + 134.2 MB\n"
                "written = 1208.0 MB (dual: 1342.2 MB).\n\n")
        kn = [short(r[name_i]) for r in data]
        f.write("| metric | unit | " + " | ".join(f"`{k}`" for k in kn) + " |\n|---|---|" + "---|" * len(kn) + "\n")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"| {w} | {units[i]} | " + " | ".join(r[i] for r in data) + " |\n")
        f.write("\nTemplate argument: <0,..> DXT1, <1,..> ETC1s, <2,..> both codecs in one pass.\n")
        # traffic per launch for bench.py's roofline.traffic
        import json
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        traffic = {}
        for r in data:
            m = re.search(r"<(\d)", r[name_i])
            key = {"0": "dxt1", "1": "etc1", "2": "dual"}.get(m.group(1) if m else "", r[name_i][:20])
            traffic[key] = float(r[ri]) * scale.get(units[ri], 1.0) + float(r[wi]) * scale.get(units[wi], 1.0)
        (out_dir / "traffic.json").write_text(json.dumps({"source": f"profiles/{tag}_ncu_full.md", "workload": "4 x 8192x8192 batched launch",
                                                          "dram_bytes_per_launch": traffic}, indent=1) + "\n")
    print("wrote", out_dir / f"{tag}_ncu_full.md")
