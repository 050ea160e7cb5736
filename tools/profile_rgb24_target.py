#!/usr/bin/env python
"""Launch sequence for one `ncu --set full` pass over the packed-RGB kernels: 4 x 8192^2 in one batched launch (the shape
tools/rgb24_ab.py and bench.py's rgb24_input leg time), every kernel twice (the second launch is the one to read)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import goofy_b200 as gb
from bench import fill_texture_device

size, n = 8192, 4
px = size * size
src = torch.empty((n, size, size, 4), dtype=torch.uint8, device="cuda")
for i in range(n):
    fill_texture_device(torch, src[i], seed=7 + i)
rgb = torch.empty((n, size, size, 3), dtype=torch.uint8, device="cuda")
rgb.copy_(src[..., :3])
del src
a = torch.empty((n, px // 2), dtype=torch.uint8, device="cuda")
b = torch.empty((n, px // 2), dtype=torch.uint8, device="cuda")
for codec in (gb.DXT1, gb.ETC1, gb.BOTH, gb.DXT1_FLOATREF, gb.ETC1_FLOATREF):
    for _ in range(2):
        gb.check(gb.encode_rgb24_device(codec, a, rgb, size, size, size * 3, d_result2=b, input_image_pitch=px * 3, result_image_pitch=px // 2, n_images=n))
    torch.cuda.synchronize()
print("done")
