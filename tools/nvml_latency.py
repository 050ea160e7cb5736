#!/usr/bin/env python
"""How long the NVML queries of bench.py's ClockSampler take, idle and while encode kernels run
(decides the sampling period that gives enough clock samples inside a ~0.2 s timed region)."""
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import pynvml as n
import torch

import goofy_b200 as gb

n.nvmlInit()
h = n.nvmlDeviceGetHandleByIndex(0)
CALLS = {
    "clock": lambda: n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM),
    "power": lambda: n.nvmlDeviceGetPowerUsage(h),
    "reasons": lambda: n.nvmlDeviceGetCurrentClocksEventReasons(h),
}


def lat(f, k=20):
    t0 = time.perf_counter()
    for _ in range(k):
        f()
    return (time.perf_counter() - t0) / k * 1e3


print("idle:", {k: round(lat(f), 3) for k, f in CALLS.items()})

size = 8192
src = torch.randint(0, 255, (4, size, size, 4), dtype=torch.uint8, device="cuda")
dst = torch.empty((4, size * size // 2), dtype=torch.uint8, device="cuda")
stop = threading.Event()
samples = []


def sampler():
    while not stop.is_set():
        t0 = time.perf_counter()
        v = [f() for f in CALLS.values()]
        samples.append((time.perf_counter() - t0) * 1e3)
        time.sleep(0.002)


for _ in range(10):
    gb.check(gb.encode_batch_uniform_device(gb.DXT1, dst, src, size, size, size * 4, size * size * 4, size * size // 2, 4))
torch.cuda.synchronize()
th = threading.Thread(target=sampler, daemon=True)
t0 = time.perf_counter()
th.start()
for _ in range(1000):
    gb.check(gb.encode_batch_uniform_device(gb.DXT1, dst, src, size, size, size * 4, size * size * 4, size * size // 2, 4))
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
stop.set()
th.join()
print(f"under load: launch loop {1e3 * (t1 - t0):.1f} ms, total {1e3 * (t2 - t0):.1f} ms, {len(samples)} samples, "
      f"per-sample ms min/median/max {min(samples):.2f}/{sorted(samples)[len(samples) // 2]:.2f}/{max(samples):.2f}")
