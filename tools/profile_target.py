#!/usr/bin/env python
"""Launch sequence for ncu: warm-up, then ONE launch each of the DXT1, ETC1s and dual-output kernels over the
batch bench.py times (4 device-resident 8192x8192 textures, one batched launch; BASELINE.json configs[1]/[2]).
    ncu --set full --clock-control none --import-source on -k regex:encode_ -s 9 -c 3 -o gpurun_out/prof \
        python tools/profile_target.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import goofy_b200 as gb
from bench import fill_texture_device

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
src = torch.empty((batch, size, size, 4), dtype=torch.uint8, device="cuda")
for b in range(batch):
    fill_texture_device(torch, src[b], seed=1 + b)
out = size * size // 2
a = torch.empty((batch, out), dtype=torch.uint8, device="cuda")
b2 = torch.empty((batch, out), dtype=torch.uint8, device="cuda")


def run_all():
    gb.check(gb.encode_batch_uniform_device(gb.DXT1, a, src, size, size, size * 4, size * size * 4, out, batch))
    gb.check(gb.encode_batch_uniform_device(gb.ETC1, a, src, size, size, size * 4, size * size * 4, out, batch))
    gb.check(gb.encode_dual_device(a, b2, src, size, size, size * 4, size * size * 4, out, batch))


for _ in range(4):
    run_all()
torch.cuda.synchronize()
print("done")
