#!/usr/bin/env python
"""What the PCIe link of this box gives a plain pinned-memory copy, beside the host API's end-to-end rate
(the e2e leg of bench.py moves 4 B/px host->device and 0.5 B/px device->host)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

import goofy_b200 as gb

size = 8192
n_in, n_out = size * size * 4, size * size // 2
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
h_in.copy_(torch.randint(0, 255, (n_in,), dtype=torch.uint8))
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ms = timed(lambda: d_in.copy_(h_in, non_blocking=True))
print(f"H2D 256 MiB pinned cudaMemcpyAsync: {ms:.3f} ms  {n_in / ms / 1e6:.1f} GB/s")
ms = timed(lambda: h_out.copy_(d_out, non_blocking=True))
print(f"D2H 32 MiB pinned cudaMemcpyAsync:  {ms:.3f} ms  {n_out / ms / 1e6:.1f} GB/s")
s2 = torch.cuda.Stream()


def both():
    d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


ms = timed(both)
print(f"H2D 256 MiB + D2H 32 MiB concurrently: {ms:.3f} ms  -> {size * size / ms / 1e3:.0f} MP/s if the encode were free")
img = h_in.view(size, size, 4)
ms = timed(lambda: gb.check(gb.compressDXT1(h_out, h_in, size, size, size * 4)))
print(f"goofy_b200.compressDXT1 on the same pinned buffers: {ms:.3f} ms  {size * size / ms / 1e3:.0f} MP/s")
