#!/usr/bin/env python
"""Is a 22 us launch limited by how fast Python can issue it?  The 16384 x 2048 strip of BASELINE.json configs[4]
(one rank's share at 8 GPUs) issued (a) through goofy_b200.encode_device (torch tensor -> pointer and stream lookups per
call), (b) through the raw ctypes entry point with integer pointers, (c) as a CUDA graph of 20 launches replayed.
Compare with tools/shapebench (C++ loop) on the same box."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import goofy_b200 as gb
from bench import fill_texture_device
from goofy_b200 import _lib

w, h, stride = 16384, 2048, 16384 * 4 + 256
buf = torch.full((h, stride), 0xAB, dtype=torch.uint8, device="cuda")
tex = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda")
fill_texture_device(torch, tex, seed=3)
buf[:, : w * 4] = tex.view(h, w * 4)
dst = torch.empty(w * h // 2, dtype=torch.uint8, device="cuda")
lib = _lib.load()
st = int(torch.cuda.current_stream().cuda_stream)
pb, pd = int(buf.data_ptr()), int(dst.data_ptr())
algo = w * h * 4.5


def timed(fn, iters):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    issue_us = (time.perf_counter() - t0) / iters * 1e6
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, issue_us


for rep in range(3):
    ms, iss = timed(lambda: gb.check(gb.encode_device(gb.DXT1, dst, buf, w, h, stride)), 200)
    print(f"api call      : {ms * 1e3:6.2f} us per launch  {algo / ms / 1e6:6.0f} GB/s   (host issues one call in {iss:.2f} us)")
    ms, iss = timed(lambda: gb.check(lib.goofy_b200_encode_device(0, pd, pb, w, h, stride, st)), 200)
    print(f"raw ctypes    : {ms * 1e3:6.2f} us per launch  {algo / ms / 1e6:6.0f} GB/s   (host issues one call in {iss:.2f} us)")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    g = torch.cuda.CUDAGraph()
    st2 = int(s.cuda_stream)
    gb.check(lib.goofy_b200_encode_device(0, pd, pb, w, h, stride, st2))
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        for _ in range(20):
            gb.check(lib.goofy_b200_encode_device(0, pd, pb, w, h, stride, int(torch.cuda.current_stream().cuda_stream)))
    for rep in range(3):
        ms, iss = timed(g.replay, 10)
        print(f"graph of 20   : {ms / 20 * 1e3:6.2f} us per launch  {algo / (ms / 20) / 1e6:6.0f} GB/s")
