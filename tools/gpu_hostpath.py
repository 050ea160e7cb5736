#!/usr/bin/env python
"""Host-pointer entry point: pinned vs pageable buffers, 8192x8192 DXT1 (diagnostic)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import goofy_b200 as gb
size = 8192
img = np.random.default_rng(0).integers(0, 256, size=size * size * 4, dtype=np.uint8)
raw = np.empty(img.size + 64, dtype=np.uint8); off = (-raw.ctypes.data) % 64
pageable = raw[off:off + img.size]; pageable[:] = img
out_pageable = np.zeros(size * size // 2, dtype=np.uint8)
pin_in = torch.empty(img.size, dtype=torch.uint8).pin_memory(); pin_in.numpy()[:] = img
pin_out = torch.empty(size * size // 2, dtype=torch.uint8).pin_memory()
def bench(name, dst, src):
    for _ in range(2): gb.check(gb.compressDXT1(dst, src, size, size, size * 4))
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); gb.check(gb.compressDXT1(dst, src, size, size, size * 4)); best = min(best, time.perf_counter() - t0)
    print(f"{name:28s} {best*1e3:8.2f} ms  {size*size/best/1e6:9.0f} MP/s")
bench("pinned in, pinned out", pin_out, pin_in)
bench("pageable in, pageable out", out_pageable, pageable)
bench("pageable in, pinned out", pin_out, pageable)
assert np.array_equal(out_pageable, pin_out.numpy())
