#!/usr/bin/env python
"""What the PCIe links of this box give plain pinned-memory copies when 1, 2, 4 or 8 ranks copy at the same time --
the floor under the end-to-end (host buffers in, host buffers out) leg of bench.py, which moves 4 B/px host->device
and 0.5 B/px device->host per step.

    python tools/h2d_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/h2d_probe.py [--no-bind]

Every rank times, with all ranks copying concurrently (barrier on both sides, CUDA events): H2D of one 8192^2 texture
(256 MiB), D2H of its blocks (32 MiB), both at once, and then the library's host call on the same pinned buffers.
Rank 0 prints one line per rank plus the max (what bench.py's max-over-ranks timing sees).  --no-bind skips binding the
process to the CPUs NVML reports as local to its GPU (bench.py binds), to show what the binding is worth."""
import argparse
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import goofy_b200 as gb
from bench import bind_to_gpu_numa_node, dist_env

ap = argparse.ArgumentParser()
ap.add_argument("--no-bind", action="store_true")
ap.add_argument("--size", type=int, default=8192)
args = ap.parse_args()

rank, local_rank, world = dist_env()
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
binding = "not bound (--no-bind)" if args.no_bind else bind_to_gpu_numa_node(local_rank)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)

size = args.size
n_in, n_out = size * size * 4, size * size // 2
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
h_in.copy_(torch.randint(0, 255, (n_in,), dtype=torch.uint8))
d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
s2 = torch.cuda.Stream(device=dev)


def barrier():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, iters=8):
    for _ in range(3):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / iters


def both():
    d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


res = [timed(lambda: d_in.copy_(h_in, non_blocking=True)), timed(lambda: h_out.copy_(d_out, non_blocking=True)), timed(both),
       timed(lambda: gb.check(gb.compressDXT1(h_out, h_in, size, size, size * 4)))]
rows = [None] * world
if dist is not None:
    dist.all_gather_object(rows, (rank, binding, res))
else:
    rows = [(rank, binding, res)]
if rank == 0:
    print(f"{world} rank(s) copying at the same time; {size}x{size} RGBA8: {n_in >> 20} MiB in, {n_out >> 20} MiB out; host cpus {os.cpu_count()}")
    print(f"{'rank':>4} {'H2D ms':>8} {'GB/s':>6} {'D2H ms':>8} {'GB/s':>6} {'both ms':>8} {'MP/s if encode were free':>24} {'host call ms':>12} {'MP/s':>7}  cpu binding")
    for r, b, (a, c, d, e) in sorted(rows):
        print(f"{r:>4} {a:8.3f} {n_in / a / 1e6:6.1f} {c:8.3f} {n_out / c / 1e6:6.1f} {d:8.3f} {size * size / d / 1e3:24.0f} {e:12.3f} {size * size / e / 1e3:7.0f}  {b}")
    worst = [max(x[2][i] for x in rows) for i in range(4)]
    print(f" max {worst[0]:8.3f} {n_in / worst[0] / 1e6:6.1f} {worst[1]:8.3f} {n_out / worst[1] / 1e6:6.1f} {worst[2]:8.3f} {size * size / worst[2] / 1e3:24.0f} "
          f"{worst[3]:12.3f} {size * size / worst[3] / 1e3:7.0f}  -> whole job {world * size * size / worst[3] / 1e3:.0f} MP/s through the host call, "
          f"{world * size * size / worst[2] / 1e3:.0f} MP/s bare copies")
if dist is not None:
    dist.destroy_process_group()
