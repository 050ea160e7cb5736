#!/usr/bin/env python
"""bench.py -- throughput of the DXT1 / ETC1s block encoders on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--codec dxt1|etc1]

Workload (BASELINE.json configs[1]): synthetic 8192x8192 RGBA8 textures, device-resident.
One "step" = one pass of the encoder over a batch of `--batch` distinct textures (default 4,
1 GiB of input, so every launch streams far more than the 126 MB L2; inputs rotate).  With N
GPUs every rank encodes its own batch (weak scaling, no data-path collective: blocks are
independent); `value` = pixels all ranks encoded / max-over-ranks device time.

One JSON line is printed by rank 0; see DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BYTES_PER_PIXEL = 4.5  # algorithmic: 4 B RGBA read + 0.5 B block written
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--codec", choices=["dxt1", "etc1"], default="dxt1")
    ap.add_argument("--size", type=int, default=8192, help="texture width = height")
    ap.add_argument("--batch", type=int, default=4, help="distinct textures per step")
    ap.add_argument("--load-path", choices=["auto", "direct", "tma", "oneshot", "async"], default="auto",
                    help="image load layer of the device entry points (see include/goofy_b200.h)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def metric_name(args) -> str:
    """ONE metric string for both arms (the driver compares them literally); where the pixels live
    (device-resident / host-resident) is stated in config.workload only."""
    return f"MP/s {args.codec.upper()} encode, {args.size}x{args.size} RGBA8"


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU while the timed region runs.
    NVML (nvidia_ml_py) every 2 ms in a thread -- the timed region of this bench lasts tens of
    milliseconds, too short for `nvidia-smi -lms 200`; nvidia-smi is the fallback."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None
        self.handle = None
        self.sm_max = None
        self._last = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except (ValueError, IndexError):
                    idx = self.gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.thread = threading.Thread(target=self._run_nvml if self.nvml else self._run_smi, daemon=True)
        self.thread.start()

    def _sample_nvml(self, with_power=True):
        n = self.nvml
        try:
            sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            if with_power or self._last is None:
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                try:
                    rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self._last = (pw, rs)
            self.samples.append((sm, self._last[0], self._last[1]))
        except Exception:
            pass

    def _run_nvml(self):
        # SM clock every 2 ms; power and throttle reasons (slower queries on some drivers) every 8th sample
        i = 0
        while not self.stop_flag.is_set():
            self._sample_nvml(with_power=(i % 8 == 0))
            i += 1
            time.sleep(0.002)

    def _run_smi(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.sm_max = float(out[1])
                self.samples.append((float(out[0]), float(out[2]), int(out[3].strip(), 16)))
            except Exception:
                return

    def stop(self) -> dict:
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=6)
        if self.nvml and len(self.samples) < 2:
            self._sample_nvml()   # very short timed regions: make sure there is at least one reading
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"]}
        sm = [x[0] for x in self.samples]
        bits = 0
        for x in self.samples:
            bits |= x[2]
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": self.sm_max,
                "power_w_max": float(max(x[1] for x in self.samples)), "samples": len(sm),
                "source": "nvml" if self.nvml else "nvidia-smi",
                "reasons": [name for name, bit in self.REASONS if bits & bit]}


def bind_to_gpu_numa_node(local_rank: int):
    """One process per GPU: run on (and first-touch pinned host buffers on) the CPUs NVML reports as local to
    this rank's GPU, so the host<->device copies of the e2e leg do not cross sockets.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = local_rank
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            idx = int(vis.split(",")[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to gpu {idx}"
    except Exception as e:  # containers often hide the topology: keep the default affinity
        return f"unbound ({type(e).__name__})"
    return "unbound"


# ----------------------------------------------------------------------------- synthetic input
def fill_texture_device(torch, t, seed: int):
    """Photo-like family S1 of SURVEY.md 8(d) (gradient + 4-bit noise), generated on the device:
    byte(x, y, c) = ((x + y) / 8 + (rnd & 15) + 20 c) & 255, rnd from torch's Philox generator."""
    h, w, _ = t.shape
    g = torch.Generator(device=t.device)
    g.manual_seed(seed)
    yy = torch.arange(h, device=t.device, dtype=torch.int32).view(h, 1, 1)
    xx = torch.arange(w, device=t.device, dtype=torch.int32).view(1, w, 1)
    cc = (torch.arange(4, device=t.device, dtype=torch.int32) * 20).view(1, 1, 4)
    rows = 1024
    for y0 in range(0, h, rows):
        y1 = min(h, y0 + rows)
        noise = torch.randint(0, 16, (y1 - y0, w, 4), device=t.device, dtype=torch.int32, generator=g)
        t[y0:y1] = (((xx + yy[y0:y1]) // 8 + noise + cc) & 255).to(torch.uint8)


# ----------------------------------------------------------------------------- reference (CPU) arm
def cpu_encode_rate(codec: int, size: int, threads: int, iters: int, ref, sample_rows: int | None = None):
    """Best-of-`iters` MP/s of the reference encoder (row-parallel over `threads`) on a size x rows sample."""
    from oracle.oracle import aligned_empty, synth_family

    rows = sample_rows or size
    img = aligned_empty(size * rows * 4)
    tile = synth_family(1, size, min(rows, 256))
    reps = (rows + tile.shape[0] - 1) // tile.shape[0]
    img[:] = np.tile(tile.reshape(-1), reps)[: img.size]
    out = np.zeros(size * rows // 2, dtype=np.uint8)
    best = float("inf")
    times = []
    for _ in range(iters + 1):
        t0 = time.perf_counter()
        rc, _ = ref.compress_mt(codec, img, size, rows, size * 4, threads, out=out)
        dt = time.perf_counter() - t0
        assert rc == 0
        times.append(dt)
        best = min(best, dt)
    return size * rows / best / 1e6, float(np.median(times[1:])) if len(times) > 1 else best


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref, built from the unmodified
    reference sources) on this box's host cores, row-parallel over all hardware threads."""
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    from oracle.oracle import DXT1, ETC1, Oracle, Reference

    codec = DXT1 if args.codec == "dxt1" else ETC1
    size = args.size
    if Reference.available():
        ref, kind = Reference(), "reference"
        threads = ref.hardware_threads() or os.cpu_count() or 1
    else:  # the oracle port, single thread
        ref, kind, threads = None, "port", 1

    from oracle.oracle import aligned_empty, synth_family
    img = aligned_empty(size * size * 4)
    tile = synth_family(1, size, 256)
    img[:] = np.tile(tile.reshape(-1), size // 256)
    out = np.zeros(size * size // 2, dtype=np.uint8)

    def one_step():
        if ref is not None:
            rc, _ = ref.compress_mt(codec, img, size, size, size * 4, threads, out=out)
        else:
            rc, _ = Oracle().compress(codec, img, size, size)
        assert rc == 0

    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = time.perf_counter() - t0
    mps = size * size * args.steps / dt / 1e6
    sample = f"{args.steps} x one {size}x{size} RGBA8 texture ({args.codec}), row-parallel over {threads} threads"
    line = {
        "impl": "reference",
        "metric": metric_name(args),
        "value": mps, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"synthetic {size}x{size} RGBA8 {args.codec.upper()}, host-resident, CPU reference",
                   "codec": args.codec, "texture": [size, size]},
        "cpu_baseline": {"value": mps, "unit": "MP/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": mps, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- B200 arm
def run_b200_arm(args):
    import torch

    import goofy_b200 as gb

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    numa = bind_to_gpu_numa_node(local_rank) if distributed else None
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    gb.set_load_path({"auto": gb.LOAD_AUTO, "direct": gb.LOAD_DIRECT, "tma": gb.LOAD_TMA, "oneshot": gb.LOAD_ONESHOT, "async": gb.LOAD_ASYNC}[args.load_path])
    codec = gb.DXT1 if args.codec == "dxt1" else gb.ETC1
    size, batch = args.size, args.batch
    stride = size * 4
    out_bytes = size * size // 2
    px_per_step = size * size * batch

    # device-resident inputs: `batch` distinct textures per rank, contiguous (uniform-batch layout)
    src = torch.empty((batch, size, size, 4), dtype=torch.uint8, device=dev)
    dst = torch.empty((batch, out_bytes), dtype=torch.uint8, device=dev)
    for b in range(batch):
        fill_texture_device(torch, src[b], seed=1000 * rank + b)
    torch.cuda.synchronize()

    img_bytes = size * size * 4

    def step(c=codec):
        # one pass over the batch through the batched device-resident entry point (one launch)
        gb.check(gb.encode_batch_uniform_device(c, dst, src, size, size, stride, img_bytes, out_bytes, batch))

    def step_per_texture(c=codec):
        # the same batch as one call (and one launch) per 8192^2 texture
        for b in range(batch):
            gb.check(gb.encode_device(c, dst[b], src[b], size, size, stride))

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        barrier()
        if sampler is not None:
            sampler.start()      # clocks are sampled during the timed region only
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = gb.kernel_launches()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = gb.kernel_launches() - l0
        if distributed:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches = timed(step, args.steps, args.warmup, sampler)
    clocks = sampler.stop() if rank == 0 else None

    total_px = px_per_step * args.steps * world
    value = total_px / (ms * 1e-3) / 1e6
    launches_per_step = launches / args.steps
    launch_ms = ms / launches
    bytes_per_launch = px_per_step * BYTES_PER_PIXEL / launches_per_step
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    peak, peak_src = hbm_peak()
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed ncu capture
    try:
        tj = json.loads((ROOT / "profiles" / "traffic.json").read_text())
        if size == 8192 and batch == 4:
            traffic = tj["dram_bytes_per_launch"].get(args.codec)
    except Exception:
        traffic = None

    # same batch, one launch per texture (what a caller encoding single 8192^2 textures sees)
    side_steps = max(min(args.steps // 2, 100), 3)
    ms_s, launches_s = timed(step_per_texture, side_steps, 3)
    value_s = px_per_step * side_steps * world / (ms_s * 1e-3) / 1e6

    # the other codec, same protocol (BASELINE.json's metric names both)
    other = gb.ETC1 if codec == gb.DXT1 else gb.DXT1
    side_steps = max(min(args.steps // 2, 100), 3)
    ms_o, _ = timed(lambda: step(other), side_steps, 3)
    value_o = px_per_step * side_steps * world / (ms_o * 1e-3) / 1e6
    achieved_o = value_o / world * 1e6 * BYTES_PER_PIXEL / 1e9

    # dual-output pass: both codecs from one read (5 B/px)
    dst2 = torch.empty((batch, out_bytes), dtype=torch.uint8, device=dev)

    def dual_step():
        gb.check(gb.encode_dual_device(dst, dst2, src, size, size, stride, img_bytes, out_bytes, batch))
    ms_d, _ = timed(dual_step, side_steps, 3)
    value_d = px_per_step * side_steps * world / (ms_d * 1e-3) / 1e6

    # ---- end to end through the drop-in host API: pinned host buffers, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        h_src = torch.empty((size, size, 4), dtype=torch.uint8).pin_memory()
        h_dst = torch.empty((out_bytes,), dtype=torch.uint8).pin_memory()
        h_src.copy_(src[0].cpu())
        host_fn = gb.compressDXT1 if codec == gb.DXT1 else gb.compressETC1

        def e2e_step():
            gb.check(host_fn(h_dst, h_src, size, size, stride))
        e2e_steps = max(3, min(args.steps, 10))
        ms_e, _ = timed(e2e_step, e2e_steps, 3)
        e2e = {"value": size * size * e2e_steps * world / (ms_e * 1e-3) / 1e6, "unit": "MP/s",
               "h2d_bytes_per_step": size * size * 4, "d2h_bytes_per_step": out_bytes,
               "ms_per_step": ms_e / e2e_steps,
               "api": f"goofy_b200.compress{args.codec.upper()}(result, input, w, h, stride) on pinned host buffers"}
        # both codecs from one upload (goofy_b200_encode_dual_host): 4 B/px in, 1 B/px out
        h_dual = torch.empty((2, out_bytes), dtype=torch.uint8).pin_memory()

        def e2e_dual_step():
            gb.check(gb.encode_dual_host(h_dual[0], h_dual[1], h_src, size, size, stride))
        ms_d2, _ = timed(e2e_dual_step, e2e_steps, 2)
        e2e["dual_output_host_call"] = {"value": size * size * e2e_steps * world / (ms_d2 * 1e-3) / 1e6,
                                        "unit": "MP/s (each pixel to DXT1 AND ETC1s, one upload)", "ms_per_step": ms_d2 / e2e_steps,
                                        "d2h_bytes_per_step": 2 * out_bytes}
        # same call with ordinary (pageable) numpy buffers: the library stages them through pinned strips
        p_src = src[0].cpu().numpy().reshape(-1)
        p_dst = np.zeros(out_bytes, dtype=np.uint8)

        def e2e_pageable_step():
            gb.check(host_fn(p_dst, p_src, size, size, stride))
        ms_p, _ = timed(e2e_pageable_step, e2e_steps, 2)
        e2e["pageable_buffers"] = {"value": size * size * e2e_steps * world / (ms_p * 1e-3) / 1e6, "unit": "MP/s",
                                   "ms_per_step": ms_p / e2e_steps}
        # the result must be the same bytes the device-resident path produced
        gb.check(gb.encode_device(codec, dst[0], src[0], size, size, stride))
        torch.cuda.synchronize()
        e2e["matches_device_path"] = bool(torch.equal(dst[0].cpu(), h_dst))

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only): the unmodified reference
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.oracle import Oracle, Reference
        if Reference.available():
            ref = Reference()
            T = ref.hardware_threads() or os.cpu_count() or 1
            one, _ = cpu_encode_rate(codec, size, 1, 6, ref)
            allc, _ = cpu_encode_rate(codec, size, T, 8, ref)
            cpu = {"value": allc, "unit": "MP/s", "cores": T, "kind": "reference",
                   "sample": f"one {size}x{size} texture, best of 8, goofy::compress{args.codec.upper()} (-O2 -msse2) "
                             f"row-parallel over {T} threads",
                   "single_thread": {"value": one, "cores": 1, "sample": f"same texture, best of 6, as shipped"}}
        else:
            o = Oracle()
            img = src[0, :1024].cpu().numpy()
            t0 = time.perf_counter()
            o.compress(codec, img, size, 1024)
            dtc = time.perf_counter() - t0
            cpu = {"value": size * 1024 / dtc / 1e6, "unit": "MP/s", "cores": 1, "kind": "port",
                   "sample": f"{size}x1024 strip, one pass of the scalar oracle"}

    if rank == 0:
        line = {
            "metric": metric_name(args),
            "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"synthetic {size}x{size} RGBA8 {args.codec.upper()}, device-resident "
                                   f"(BASELINE.json configs[{1 if args.codec == 'dxt1' else 2}])",
                       "codec": args.codec, "texture": [size, size], "textures_per_step": batch,
                       "pixels_per_step_per_gpu": px_per_step, "stride": stride,
                       "l2": f"inputs larger than L2: every step streams {batch} distinct {size * size * 4 >> 20} MiB textures "
                             f"({int(px_per_step * BYTES_PER_PIXEL) >> 20} MiB per step vs 126 MB of L2)",
                       "load_path": args.load_path,
                       "sharding": "one batch per rank, no collectives" if world > 1 else "single GPU",
                       "cpu_binding": numa},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                         "kernel": ("encode_direct_kernel" if codec == gb.DXT1 and args.load_path in ("auto", "oneshot")
                                    else "encode_tma_kernel" if args.load_path == "tma" else "encode_rows_kernel") + f"<{args.codec}>",
                         "bytes_per_launch": bytes_per_launch, "launch_ms": launch_ms,
                         "like_for_like_ceiling": "tools/membench (profiles/r01_membench.txt): a trivial-compute kernel with the same "
                                                  "8:1 read:write access pattern reaches 7115 GB/s at 1.2 GB per launch "
                                                  "(6696 GB/s at 302 MB); read-only 7355 GB/s"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "per_texture_launch": {"value": value_s, "unit": "MP/s", "launches_per_step": launches_s / side_steps,
                                   "achieved_gbs_per_gpu": value_s / world * 1e6 * BYTES_PER_PIXEL / 1e9,
                                   "note": "same batch, one goofy_b200_encode_device call per texture"},
            "other_codec": {"codec": "etc1" if codec == gb.DXT1 else "dxt1", "value": value_o, "unit": "MP/s",
                            "achieved_gbs_per_gpu": achieved_o, "frac": achieved_o / peak},
            "dual_output": {"value": value_d, "unit": "MP/s (each pixel encoded to both DXT1 and ETC1s)",
                            "achieved_gbs_per_gpu": value_d / world * 1e6 * 5.0 / 1e9},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    # Keep stdout to the one JSON line: libraries (NCCL prints its version banner there) get stderr.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")
    real_print = print

    def print_json(*a, **k):
        real_print(*a, **k, file=json_out, flush=True)
    globals()["print"] = print_json
    try:
        if args.impl == "reference":
            return run_reference_arm(args)
        return run_b200_arm(args)
    finally:
        globals()["print"] = real_print
        json_out.flush()


if __name__ == "__main__":
    sys.exit(main())
