#!/usr/bin/env python
"""Throughput of the packed-RGB kernels (goofy_b200_encode_rgb24_device) on 4 x 8192^2 device-resident textures, one
batched launch per step like bench.py's headline.  The library is whatever GOOFY_B200_LIB points at (experiments).

    [GOOFY_B200_LIB=build/ab/libgoofy_X.so] [GOOFY_B200_RGB24_ROWS_PER_CTA=n] python tools/rgb24_ab.py [--steps 100]
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import goofy_b200 as gb
from bench import fill_texture_device

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--label", default="")
args = ap.parse_args()
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
res = []
for name, codec, bpp in (("dxt1", gb.DXT1, 3.5), ("etc1s", gb.ETC1, 3.5), ("both", gb.BOTH, 4.0)):
    def fn():
        gb.check(gb.encode_rgb24_device(codec, a, rgb, size, size, size * 3, d_result2=b, input_image_pitch=px * 3, result_image_pitch=px // 2, n_images=n))
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    res.append(f"{name} {n * px * bpp / ms / 1e6:6.0f} GB/s {n * px / ms / 1e9:5.3f} TP/s")
print(f"{args.label:28s} " + " | ".join(res))
