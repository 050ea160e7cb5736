#!/usr/bin/env python
"""Host-pointer entry point: pinned vs pageable buffers, DXT1 (diagnostic).  usage: gpu_hostpath.py [width height]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import goofy_b200 as gb
W = int(sys.argv[1]) if len(sys.argv) > 2 else 8192
H = int(sys.argv[2]) if len(sys.argv) > 2 else W
img = np.random.default_rng(0).integers(0, 256, size=W * H * 4, dtype=np.uint8)
raw = np.empty(img.size + 64, dtype=np.uint8); off = (-raw.ctypes.data) % 64
pageable = raw[off:off + img.size]; pageable[:] = img
out_pageable = np.zeros(W * H // 2, dtype=np.uint8)
pin_in = torch.empty(img.size, dtype=torch.uint8).pin_memory(); pin_in.numpy()[:] = img
pin_out = torch.empty(W * H // 2, dtype=torch.uint8).pin_memory()
def bench(name, dst, src):
    for _ in range(6): gb.check(gb.compressDXT1(dst, src, W, H, W * 4))
    ts = []
    for _ in range(24 if W * H > (1 << 22) else 200):
        t0 = time.perf_counter(); gb.check(gb.compressDXT1(dst, src, W, H, W * 4)); ts.append(time.perf_counter() - t0)
    best, med = min(ts), sorted(ts)[len(ts) // 2]
    print(f"{W}x{H} {name:28s} best {best*1e6:9.1f} us  median {med*1e6:9.1f} us  {W*H/best/1e6:9.0f} MP/s")
bench("pinned in, pinned out", pin_out, pin_in)
bench("pageable in, pageable out", out_pageable, pageable)
bench("pageable in, pinned out", pin_out, pageable)
assert np.array_equal(out_pageable, pin_out.numpy())
