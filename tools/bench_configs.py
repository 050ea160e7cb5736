#!/usr/bin/env python
"""BASELINE.json configs[3] and configs[4] (the multi-GPU shapes), same timing rules as bench.py.

    [torchrun --nproc-per-node N ...] tools/bench_configs.py --config batch1024  [--images 4096]
    [torchrun --nproc-per-node N ...] tools/bench_configs.py --config strip16384 [--load-path tma]

batch1024   4096 textures of 1024x1024, DXT1 + ETC1s, textures partitioned across the ranks
            (goofy_b200.sharding.batch_partition), one uniform-batch launch per codec per step, plus the
            dual-output kernel (both codecs from one read).  Strong scaling: total work is fixed.
strip16384  one 16384x16384 texture with padded row stride 65792 B (pad bytes 0xAB), DXT1, strip g of N on
            rank g (goofy_b200_strip_partition).  Strong scaling.
No collectives on the data path; NCCL only for the barrier and the max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import goofy_b200 as gb  # noqa: E402
from bench import BYTES_PER_PIXEL, ClockSampler, dist_env, fill_texture_device, hbm_peak  # noqa: E402
from goofy_b200 import sharding  # noqa: E402


def main():
    import os
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")   # stdout carries the JSON line only; library banners go to stderr
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=["batch1024", "strip16384"], required=True)
    ap.add_argument("--images", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--load-path", choices=["auto", "direct", "tma", "oneshot", "async"], default="auto")
    args = ap.parse_args()

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    gb.set_load_path({"auto": gb.LOAD_AUTO, "direct": gb.LOAD_DIRECT, "tma": gb.LOAD_TMA, "oneshot": gb.LOAD_ONESHOT, "async": gb.LOAD_ASYNC}[args.load_path])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = gb.kernel_launches()
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, gb.kernel_launches() - l0

    peak, peak_src = hbm_peak()
    results = {}
    if args.config == "batch1024":
        w = h = 1024
        mine = sharding.batch_partition(args.images, world, rank)
        n = len(mine)
        # the rank's shard, generated on the device: a few distinct textures tiled over the shard
        src = torch.empty((n, h, w, 4), dtype=torch.uint8, device=dev)
        distinct = min(n, 16)
        for i in range(distinct):
            fill_texture_device(torch, src[i], seed=7 * (mine.start + i) + 1)
        for i in range(distinct, n):
            src[i].copy_(src[i % distinct])
        d1 = torch.empty((n, w * h // 2), dtype=torch.uint8, device=dev)
        d2 = torch.empty((n, w * h // 2), dtype=torch.uint8, device=dev)
        total_px = args.images * w * h

        def both():
            gb.check(gb.encode_batch_uniform_device(gb.DXT1, d1, src, w, h, w * 4, w * h * 4, w * h // 2, n))
            gb.check(gb.encode_batch_uniform_device(gb.ETC1, d2, src, w, h, w * 4, w * h * 4, w * h // 2, n))

        def dual():
            gb.check(gb.encode_dual_device(d1, d2, src, w, h, w * 4, w * h * 4, w * h // 2, n))

        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ms_b, launches = timed(both)
        clocks = sampler.stop() if rank == 0 else None
        ms_d, _ = timed(dual)
        # check the two ways agree on this rank's shard
        both()
        a1, a2 = d1.clone(), d2.clone()
        dual()
        torch.cuda.synchronize()
        same = bool(torch.equal(a1, d1) and torch.equal(a2, d2))
        results = {
            "two_passes": {"value": total_px * args.steps / (ms_b * 1e-3) / 1e6, "unit": "MP/s (each pixel to DXT1 AND ETC1s)",
                           "ms_per_step": ms_b / args.steps,
                           "achieved_gbs_per_gpu": total_px / world * 9.0 * args.steps / (ms_b * 1e-3) / 1e9, "bytes_per_pixel": 9.0},
            "dual_kernel": {"value": total_px * args.steps / (ms_d * 1e-3) / 1e6, "unit": "MP/s (each pixel to DXT1 AND ETC1s)",
                            "ms_per_step": ms_d / args.steps,
                            "achieved_gbs_per_gpu": total_px / world * 5.0 * args.steps / (ms_d * 1e-3) / 1e9, "bytes_per_pixel": 5.0},
            "dual_equals_two_passes": same,
        }
        workload = f"{args.images} x 1024x1024 RGBA8, DXT1+ETC1s, textures sharded over {world} GPU(s) (BASELINE.json configs[3])"
        value = results["dual_kernel"]["value"]
        ms_step = ms_d / args.steps
    else:
        w = h = 16384
        stride = w * 4 + 256
        first, count = sharding.strip_partition(h, world, rank)
        rows = count * 4
        buf = torch.full((rows, stride), 0xAB, dtype=torch.uint8, device=dev)
        tex = torch.empty((rows, w, 4), dtype=torch.uint8, device=dev)
        fill_texture_device(torch, tex, seed=31 + rank)
        buf[:, : w * 4] = tex.view(rows, w * 4)
        del tex
        dst = torch.empty(rows * w // 2, dtype=torch.uint8, device=dev)
        total_px = w * h

        def strip():
            gb.check(gb.encode_device(gb.DXT1, dst, buf, w, rows, stride))

        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ms, launches = timed(strip)
        clocks = sampler.stop() if rank == 0 else None
        value = total_px * args.steps / (ms * 1e-3) / 1e6
        ms_step = ms / args.steps
        results = {"achieved_gbs_per_gpu": total_px / world * BYTES_PER_PIXEL * args.steps / (ms * 1e-3) / 1e9,
                   "strip_rows_per_gpu": rows, "stride": stride}
        workload = f"16384x16384 RGBA8, stride {stride} B, DXT1, {world} strip(s) of whole block rows (BASELINE.json configs[4])"

    if rank == 0:
        print(json.dumps({
            "metric": "MP/s", "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "load_path": args.load_path}, "results": results, "hbm_peak_gbs": peak,
            "peak_source": peak_src, "gpu_launches": int(launches), "clocks": clocks}), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
