#!/usr/bin/env python
"""Launch sequence for ncu: warm-up, then ONE launch each of the DXT1, ETC1s and dual-output kernels on a
device-resident 8192x8192 texture (BASELINE.json configs[1]/[2]).  Run as
    ncu --set full --clock-control none --import-source on -k regex:encode_direct -s 9 -c 3 -o gpurun_out/prof \
        python tools/profile_target.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import goofy_b200 as gb
from bench import fill_texture_device

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
src = torch.empty((size, size, 4), dtype=torch.uint8, device="cuda")
fill_texture_device(torch, src, seed=1)
a = torch.empty(size * size // 2, dtype=torch.uint8, device="cuda")
b = torch.empty(size * size // 2, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run_all():
    flush.zero_()
    gb.check(gb.encode_device(gb.DXT1, a, src, size, size, size * 4))
    flush.zero_()
    gb.check(gb.encode_device(gb.ETC1, a, src, size, size, size * 4))
    flush.zero_()
    gb.check(gb.encode_dual_device(a, b, src, size, size, size * 4))


for _ in range(4):
    run_all()
torch.cuda.synchronize()
print("done")
