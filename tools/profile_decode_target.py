#!/usr/bin/env python
"""Launch sequence for ncu: the decoders and the fused decode + squared-error kernels on one 8192x8192 texture each
(SURVEY.md 8(f) row N2).
    ncu --set full --clock-control none --import-source on -k regex:"decode_kernel|block_sse" -s 12 -c 4 -f -o gpurun_out/prof_decode \
        python tools/profile_decode_target.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import goofy_b200 as gb
from bench import fill_texture_device

size = 8192
src = torch.empty((size, size, 4), dtype=torch.uint8, device="cuda")
fill_texture_device(torch, src, seed=3)
ob = size * size // 2
blk = {c: torch.empty(ob, dtype=torch.uint8, device="cuda") for c in (gb.DXT1, gb.ETC1)}
for c in blk:
    gb.check(gb.encode_device(c, blk[c], src, size, size, size * 4))
dec = torch.empty((size, size, 4), dtype=torch.uint8, device="cuda")
sse = torch.zeros(3, dtype=torch.int64, device="cuda")
for _ in range(4):
    for c in (gb.DXT1, gb.ETC1):
        gb.check(gb.decode_device(c, dec, blk[c], size, size, size * 4))
    for c in (gb.DXT1, gb.ETC1):
        gb.check(gb.block_sse_device(c, blk[c], src, size, size, size * 4, sse))
torch.cuda.synchronize()
print("done")
