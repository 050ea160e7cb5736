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
    ap.add_argument("--no-traffic-probe", action="store_true",
                    help="do not re-measure roofline.traffic with ncu after the timed legs (use the committed capture)")
    ap.add_argument("--no-configs", action="store_true", help="skip the named multi-GPU shapes (configs[3], configs[4], shard scheduler)")
    ap.add_argument("--images", type=int, default=4096, help="textures of configs[3] (4096 x 1024^2 in BASELINE.json)")
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


def measure_traffic_with_ncu(codec_name: str, size: int, batch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the headline shape, measured now: a child process runs
    tools/profile_target.py (the same batched launches, four rounds) under ncu, which captures the last round.  Counters
    only -- nothing timed comes from that run.  Returns (bytes, source) or (None, why)."""
    import csv
    import io
    import shutil
    # not from inside a profiler: a bench run under ncu / nsys / compute-sanitizer must not start a second profiler
    # (ncu, nsys and compute-sanitizer inject themselves through *INJECTION* variables; the image itself sets NV_CUDA_NSIGHT_*)
    injected = [k for k in os.environ if "INJECTION" in k.upper()]
    if injected:
        return None, "this run is itself under a profiler (" + injected[0] + ")"
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--print-units", "base", "--clock-control", "none",
           "-k", "regex:encode_", "-s", "9", "-c", "3", "--csv", sys.executable, str(ROOT / "tools" / "profile_target.py"), str(size), str(batch)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=str(ROOT)).stdout
    except Exception as e:  # noqa: BLE001
        return None, f"ncu failed ({type(e).__name__})"
    lines = [l for l in out.splitlines() if l.startswith('"')]
    if len(lines) < 2:
        return None, "ncu printed no counters"
    per_launch = {}
    for r in csv.DictReader(io.StringIO("\n".join(lines))):
        try:
            per_launch.setdefault(r["ID"], [r["Kernel Name"], 0.0])[1] += float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            return None, "unexpected ncu output"
    order = sorted(per_launch, key=lambda k: int(k))      # the captured round: DXT1, ETC1s, dual-output
    want = {"dxt1": 0, "etc1": 1}[codec_name]
    if len(order) != 3:
        return None, f"ncu captured {len(order)} launches, expected 3"
    name, total = per_launch[order[want]]
    return total, f"measured after the timed legs of this run: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on one launch of {name.split('(')[0]} over the same batch (tools/profile_target.py)"


# ----------------------------------------------------------------------------- clocks
_SAMPLER_CHILD = r"""
import json, sys, time
import pynvml as n
n.nvmlInit()
h = n.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
sm_max = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
print(json.dumps({"ready": True, "sm_max": sm_max}), flush=True)
import select
samples = []
i = 0
pw, rs = 0.0, 0
while True:
    if i % 8 == 0:
        try:
            pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0
            try: rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception: rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        except Exception: pass
    try: samples.append((round(time.monotonic(), 6), float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), pw, rs))
    except Exception: pass
    i += 1
    if select.select([sys.stdin], [], [], 0.00002)[0]:    # a dozen samples per millisecond
        break
    if len(samples) > 600000: break
print(json.dumps({"samples": samples}), flush=True)
"""


class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU while the timed regions run.

    The sampling loop lives in a CHILD PROCESS (NVML through nvidia_ml_py, back to back, a few samples per
    millisecond): a thread of this process only gets the interpreter lock every few milliseconds while the main thread
    is busy launching, which left 2 samples in a 3.4 ms region.  Samples carry CLOCK_MONOTONIC timestamps (system-wide),
    so the parent keeps the ones that fall between its own begin / end marks.  Fallback: nvidia-smi, then nothing."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.sm_max = None
        self.marks = {}
        idx = gpu_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                idx = int(vis.split(",")[gpu_index])
            except (ValueError, IndexError):
                idx = gpu_index
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_CHILD, str(idx)], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            ready = json.loads(self.proc.stdout.readline())
            self.sm_max = ready["sm_max"]
        except Exception:
            self.proc = None

    def begin(self, name: str):
        self.marks[name] = [time.monotonic(), None]

    def end(self, name: str):
        self.marks[name][1] = time.monotonic()

    # timed() brackets its region with these two
    def start(self):
        self.begin("headline")

    def stop_region(self):
        self.end("headline")

    def finish(self, span_name: str = "headline") -> dict:
        """Stop the child and summarise the samples inside the `span_name` marks (and count the headline's)."""
        samples = []
        if self.proc is not None:
            try:
                self.proc.stdin.write("stop\n")
                self.proc.stdin.flush()
                samples = json.loads(self.proc.stdout.readline())["samples"]
                self.proc.wait(timeout=5)
            except Exception:
                samples = []
        if not samples:
            return self._smi_once()
        t0, t1 = self.marks.get(span_name, self.marks.get("headline"))
        h0, h1 = self.marks["headline"]
        inside = [x for x in samples if t0 <= x[0] <= (t1 or x[0])]
        head = [x for x in samples if h0 <= x[0] <= (h1 or x[0])]

        def summary(use):
            sm = [x[1] for x in use]
            bits = 0
            for x in use:
                bits |= int(x[3])
            return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": self.sm_max,
                    "power_w_max": float(max(x[2] for x in use)), "samples": len(use),
                    "reasons": [name for name, bit in self.REASONS if bits & bit]}
        # the top-level figures describe the HEADLINE timed region when it holds enough samples (it lasts 3.4 ms at
        # 20 steps); `all_timed_legs` covers every timed leg of the run (end-to-end and named-config legs included)
        out = summary(head if len(head) >= 3 else (inside or samples[-3:]))
        out["region"] = "headline timed region" if len(head) >= 3 else "all timed legs (the headline region held fewer than 3 samples)"
        out["source"] = "nvml (child process, CLOCK_MONOTONIC-stamped)"
        if inside:
            out["all_timed_legs"] = summary(inside)
        return out

    def _smi_once(self) -> dict:
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active"
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                 capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            bits = int(out[3].strip(), 16)
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "power_w_max": float(out[2]), "samples": 1,
                    "source": "nvidia-smi (one reading after the timed regions: NVML sampler unavailable)",
                    "reasons": [name for name, bit in self.REASONS if bits & bit]}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"]}


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
class Ctx:
    """Everything the legs of the B200 arm share: device, rank, timing helpers."""

    def __init__(self, args):
        import torch

        import goofy_b200 as gb
        self.torch, self.gb, self.args = torch, gb, args
        self.rank, self.local_rank, self.world = dist_env()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.distributed = self.world > 1
        self.numa = bind_to_gpu_numa_node(self.local_rank) if self.distributed else None
        self.dist = None
        self.host_group = None
        if self.distributed:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
            # a second, CPU-side group: ranks that merely wait for rank 0 must not do it inside an NCCL kernel that
            # spins on their GPU (rank 0 may be using that GPU through the in-library shard scheduler)
            self.host_group = dist.new_group(backend="gloo")
        # one process per GPU on a shared host: the host path's staging threads (copy / alpha strip) are divided between
        # the local ranks; set before the library starts its pool (an explicit GOOFY_B200_HOST_THREADS wins)
        if self.distributed and "GOOFY_B200_HOST_THREADS" not in os.environ:
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", self.world))
            os.environ["GOOFY_B200_HOST_THREADS"] = str(max(2, min(8, (os.cpu_count() or 8) // max(1, local_world))))
        # (whether pinned input is alpha-stripped is the library's call: with other ranks' processes on the box's GPUs it
        # sees neighbours -- goofy_b200_host_neighbours, reported in e2e.host_neighbours -- and leaves pinned input to plain DMA)
        from goofy_b200 import _lib
        self.lib = _lib.load()          # raw ctypes entry points: pointers and the stream as plain integers, so a
        self.stream = int(torch.cuda.current_stream().cuda_stream)   # 20 us launch is not waiting for Python

    def barrier(self):
        if self.distributed:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):
        """Wait for the other ranks on the CPU (gloo), GPUs idle."""
        self.torch.cuda.synchronize()
        if self.distributed:
            self.dist.barrier(group=self.host_group)

    def all_max(self, x: float) -> float:
        if not self.distributed:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_values(self, x: float):
        """x of every rank, in rank order (on every rank)."""
        if not self.distributed:
            return [x]
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(v.item()) for v in out]

    def all_true(self, ok: bool) -> bool:
        if not self.distributed:
            return bool(ok)
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def timed(self, fn, steps, warmup, sampler=None, collective=True):
        """W warm-up steps, then K steps between two CUDA events on the launching stream, barrier + synchronize on
        both sides, MAX over ranks.  collective=False: this rank alone (the others are parked at a barrier)."""
        torch, gb = self.torch, self.gb
        for _ in range(warmup):
            fn()
        self.barrier() if collective else torch.cuda.synchronize()
        if sampler is not None:
            sampler.start()      # clocks are sampled during the timed region only
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = gb.kernel_launches()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier() if collective else torch.cuda.synchronize()
        if sampler is not None:
            sampler.stop_region()
        ms = e0.elapsed_time(e1)
        launches = gb.kernel_launches() - l0
        self.last_ms_by_rank = self.all_values(ms) if collective else [ms]
        return (max(self.last_ms_by_rank) if collective else ms), launches


def reference_blocks(codec: int, img: np.ndarray, w: int, h: int, stride: int) -> np.ndarray:
    """The checker of the bit-exactness samples: the unmodified reference (oracle/_ref) where it was built, else the
    scalar oracle.  Test infrastructure -- never on a timed path."""
    from oracle.oracle import Oracle, Reference, aligned_copy
    if Reference.available():
        ref = Reference()
        rc, out = ref.compress_mt(codec, aligned_copy(img), w, h, stride, max(1, min(8, ref.hardware_threads() or 1)))
    else:
        rc, out = Oracle().compress(codec, aligned_copy(img), w, h, stride)
    assert rc == 0
    return out


def config_batch1024(ctx: Ctx, steps: int, warmup: int, n_images: int):
    """BASELINE.json configs[3]: `n_images` textures of 1024 x 1024, DXT1 + ETC1s, partitioned by texture over the ranks
    (goofy_b200.sharding.batch_partition; GoofyTC/goofy_tc.h:1514-1524 has no cross-block state to exchange).
    Strong scaling: the total is fixed.  Timed: the dual-output kernel (one read of every pixel, 5 B/px) and the two
    single-codec passes (9 B/px); checked: a sample of this rank's shard against the reference, both codecs."""
    torch, gb = ctx.torch, ctx.gb
    from goofy_b200 import sharding
    w = h = 1024
    img_bytes, out_bytes = w * h * 4, w * h // 2

    def make_shard(indices):
        n = len(indices)
        src = torch.empty((n, h, w, 4), dtype=torch.uint8, device=ctx.dev)
        distinct = min(n, 16)
        for i in range(distinct):
            fill_texture_device(torch, src[i], seed=7 * (indices.start + i) + 1)
        if n > distinct:   # a few textures of the stress families S0 (uniform random) and S2 (0 / 255), SURVEY.md 8(d)
            g = torch.Generator(device=ctx.dev)
            g.manual_seed(1234 + indices.start)
            src[distinct - 2] = torch.randint(0, 256, (h, w, 4), device=ctx.dev, dtype=torch.int32, generator=g).to(torch.uint8)
            src[distinct - 1] = (torch.randint(0, 2, (h, w, 4), device=ctx.dev, dtype=torch.int32, generator=g) * 255).to(torch.uint8)
        for i in range(distinct, n):
            src[i].copy_(src[i % distinct])
        return src, torch.empty((n, out_bytes), dtype=torch.uint8, device=ctx.dev), torch.empty((n, out_bytes), dtype=torch.uint8, device=ctx.dev)

    def runners(src, d1, d2):
        n = src.shape[0]
        ps, p1, p2 = int(src.data_ptr()), int(d1.data_ptr()), int(d2.data_ptr())
        lib, st = ctx.lib, ctx.stream

        def dual():
            gb.check(lib.goofy_b200_encode_dual_device(p1, p2, ps, w, h, w * 4, img_bytes, out_bytes, n, st))

        def two_passes():
            gb.check(lib.goofy_b200_encode_batch_uniform_device(gb.DXT1, p1, ps, w, h, w * 4, img_bytes, out_bytes, n, st))
            gb.check(lib.goofy_b200_encode_batch_uniform_device(gb.ETC1, p2, ps, w, h, w * 4, img_bytes, out_bytes, n, st))
        return dual, two_passes

    mine = sharding.batch_partition(n_images, ctx.world, ctx.rank)
    src, d1, d2 = make_shard(mine)
    dual, two_passes = runners(src, d1, d2)
    total_px = n_images * w * h
    ms_d, launches = ctx.timed(dual, steps, warmup)
    by_rank = [m / steps for m in ctx.last_ms_by_rank]
    kernel = gb.last_launch_kernel()
    ms_t, _ = ctx.timed(two_passes, steps, warmup)
    # bit-exactness of a sample of this rank's shard: first, a stress-family and the last texture, both codecs
    dual()
    torch.cuda.synchronize()
    ok = True
    n = len(mine)
    for i in sorted({0, min(n - 1, 14), min(n - 1, 15), n - 1}):
        host = src[i].cpu().numpy()
        ok = ok and np.array_equal(d1[i].cpu().numpy(), reference_blocks(0, host, w, h, w * 4))
        ok = ok and np.array_equal(d2[i].cpu().numpy(), reference_blocks(1, host, w, h, w * 4))
    two_passes()
    a1, a2 = d1.clone(), d2.clone()
    dual()
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(a1, d1) and torch.equal(a2, d2))
    ok = ctx.all_true(ok)
    del a1, a2
    out = {
        "workload": f"{n_images} x 1024x1024 RGBA8, DXT1+ETC1s, textures partitioned over {ctx.world} GPU(s) (BASELINE.json configs[3])",
        "scaling": "strong", "sharding": "batch_partition (contiguous texture ranges, no collectives)",
        "value": total_px * steps / (ms_d * 1e-3) / 1e6, "unit": "MP/s (each pixel to DXT1 AND ETC1s)", "ms_per_step": ms_d / steps,
        "kernel": kernel, "bytes_per_pixel": 5.0, "achieved_gbs_per_gpu": total_px / ctx.world * 5.0 * steps / (ms_d * 1e-3) / 1e9,
        "two_passes": {"value": total_px * steps / (ms_t * 1e-3) / 1e6, "bytes_per_pixel": 9.0,
                       "achieved_gbs_per_gpu": total_px / ctx.world * 9.0 * steps / (ms_t * 1e-3) / 1e9},
        "bit_exact": ok, "bit_exact_sample": "4 textures of every rank's shard vs the reference (both codecs), and dual == two passes on the whole shard",
        "steps": steps, "gpu_launches": int(launches), "ms_per_step_by_rank": by_rank,
    }
    if ctx.world > 1:
        # the same workload on ONE GPU of this box (rank 0, the others parked at the barrier): the denominator of the efficiency
        del src, d1, d2
        torch.cuda.empty_cache()
        v1 = None
        if ctx.rank == 0:
            src, d1, d2 = make_shard(range(n_images))
            dual1, _ = runners(src, d1, d2)
            ms1, _ = ctx.timed(dual1, max(3, steps // 2), 2, collective=False)
            v1 = total_px * max(3, steps // 2) / (ms1 * 1e-3) / 1e6
            del src, d1, d2
            torch.cuda.empty_cache()
        ctx.host_barrier()
        if ctx.rank == 0:
            out["n1_value_same_box"] = v1
            out["efficiency_vs_n1_per_gpu"] = out["value"] / (ctx.world * v1)
    return out


def config_strip16384(ctx: Ctx, steps: int, warmup: int):
    """BASELINE.json configs[4]: one 16384 x 16384 texture whose rows are 65 792 bytes apart (pad bytes 0xAB), DXT1,
    strip g of N on rank g (goofy_b200_strip_partition).  Strong scaling.  Checked: the first and the last 64 rows of
    this rank's strip against the reference run on the same padded rows."""
    torch, gb = ctx.torch, ctx.gb
    w = h = 16384
    stride = w * 4 + 256

    def make_strip(first, count, seed):
        rows = count * 4
        buf = torch.full((rows, stride), 0xAB, dtype=torch.uint8, device=ctx.dev)
        for y0 in range(0, rows, 2048):
            y1 = min(rows, y0 + 2048)
            tex = torch.empty((y1 - y0, w, 4), dtype=torch.uint8, device=ctx.dev)
            fill_texture_device(torch, tex, seed=seed + first * 4 + y0)
            buf[y0:y1, : w * 4] = tex.view(y1 - y0, w * 4)
            del tex
        return buf, torch.empty(rows * w // 2, dtype=torch.uint8, device=ctx.dev)

    def runner(buf, dst):
        pb, pd, rows = int(buf.data_ptr()), int(dst.data_ptr()), buf.shape[0]
        lib, st = ctx.lib, ctx.stream
        return lambda: gb.check(lib.goofy_b200_encode_device(gb.DXT1, pd, pb, w, rows, stride, st))

    first, count = gb.strip_partition(h, ctx.world, ctx.rank)
    buf, dst = make_strip(first, count, 31)
    strip = runner(buf, dst)
    ms, launches = ctx.timed(strip, steps, warmup)
    by_rank = [m / steps for m in ctx.last_ms_by_rank]
    kernel = gb.last_launch_kernel()
    strip()
    torch.cuda.synchronize()
    ok = True
    rows = count * 4
    for y0 in sorted({0, rows - 64}):
        host = buf[y0:y0 + 64].cpu().numpy()
        want = reference_blocks(0, host, w, 64, stride)
        got = dst[(y0 // 4) * (w // 4) * 8: (y0 // 4 + 16) * (w // 4) * 8].cpu().numpy()
        ok = ok and np.array_equal(got, want)
    ok = ctx.all_true(ok)
    total_px = w * h
    out = {
        "workload": f"16384x16384 RGBA8, stride {stride} B (pad 0xAB), DXT1, {ctx.world} strip(s) of whole block rows (BASELINE.json configs[4])",
        "scaling": "strong", "sharding": "strip_partition (rows [2048g, 2048(g+1)) at 8 GPUs, no collectives)",
        "value": total_px * steps / (ms * 1e-3) / 1e6, "unit": "MP/s", "ms_per_step": ms / steps, "kernel": kernel,
        "bytes_per_pixel": BYTES_PER_PIXEL, "achieved_gbs_per_gpu": total_px / ctx.world * BYTES_PER_PIXEL * steps / (ms * 1e-3) / 1e9,
        "strip_rows_per_gpu": rows, "stride": stride,
        "bit_exact": ok, "bit_exact_sample": "first and last 64 pixel rows of every rank's strip vs the reference on the same padded rows",
        "steps": steps, "gpu_launches": int(launches), "ms_per_step_by_rank": by_rank,
    }
    if ctx.world > 1:
        del buf, dst
        torch.cuda.empty_cache()
        v1 = None
        if ctx.rank == 0:
            buf, dst = make_strip(0, h // 4, 31)
            whole = runner(buf, dst)
            k = max(3, steps // 2)
            ms1, _ = ctx.timed(whole, k, 2, collective=False)
            v1 = total_px * k / (ms1 * 1e-3) / 1e6
            del buf, dst
            torch.cuda.empty_cache()
        ctx.host_barrier()
        if ctx.rank == 0:
            out["n1_value_same_box"] = v1
            out["efficiency_vs_n1_per_gpu"] = out["value"] / (ctx.world * v1)
    return out


def sharded_api_leg(ctx: Ctx):
    """The in-library scheduler (one host thread per device, DeviceWorker in host_batch.cuh) across ALL devices this
    process can see, called by rank 0 alone while the other ranks wait: goofy_b200_encode_batch_sharded on textures
    resident on every device (both codecs from one read) and goofy_b200_encode_sharded_host on one host image in strips.
    Both checked against the reference."""
    torch, gb = ctx.torch, ctx.gb
    out = None
    ctx.host_barrier()     # every rank's GPU is idle from here until the closing host barrier
    if ctx.rank == 0:
        n_dev = gb.device_count()
        w = h = 1024
        per_dev = 8
        keep, descs = [], []
        for d in range(n_dev):
            with torch.cuda.device(d):
                for k in range(per_dev):
                    s = torch.empty((h, w, 4), dtype=torch.uint8, device=f"cuda:{d}")
                    fill_texture_device(torch, s, seed=900 + per_dev * d + k)
                    a = torch.zeros(w * h // 2, dtype=torch.uint8, device=f"cuda:{d}")
                    b = torch.zeros(w * h // 2, dtype=torch.uint8, device=f"cuda:{d}")
                    keep.append((s, a, b))
                    descs.append((s, a, w, h, w * 4, d, b))
                torch.cuda.synchronize(d)
        arr = gb.make_descriptors(descs)
        gb.check(gb.encode_batch_sharded(gb.BOTH, arr))
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            gb.check(gb.encode_batch_sharded(gb.BOTH, arr))
        dt = (time.perf_counter() - t0) / reps
        ok = True
        for i in sorted({0, len(keep) // 2, len(keep) - 1}):
            s, a, b = keep[i]
            host = s.cpu().numpy()
            ok = ok and np.array_equal(a.cpu().numpy(), reference_blocks(0, host, w, h, w * 4))
            ok = ok and np.array_equal(b.cpu().numpy(), reference_blocks(1, host, w, h, w * 4))
        # one host image (pinned), strips of whole block rows, strip g on device g
        hw, hh = 8192, 2048
        h_src = torch.empty((hh, hw, 4), dtype=torch.uint8).pin_memory()
        h_dst = torch.empty(hw * hh // 2, dtype=torch.uint8).pin_memory()
        with torch.cuda.device(0):
            tmp = torch.empty((hh, hw, 4), dtype=torch.uint8, device="cuda:0")
            fill_texture_device(torch, tmp, seed=77)
            h_src.copy_(tmp.cpu())
            del tmp
        gb.check(gb.encode_sharded_host(gb.DXT1, h_dst, h_src, hw, hh, hw * 4, n_dev))
        t0 = time.perf_counter()
        for _ in range(reps):
            gb.check(gb.encode_sharded_host(gb.DXT1, h_dst, h_src, hw, hh, hw * 4, n_dev))
        dth = (time.perf_counter() - t0) / reps
        band = 256
        for y0 in (0, hh - band):
            want = reference_blocks(0, h_src[y0:y0 + band].numpy(), hw, band, hw * 4)
            ok = ok and np.array_equal(h_dst[(y0 // 4) * (hw // 4) * 8: ((y0 + band) // 4) * (hw // 4) * 8].numpy(), want)
        out = {"n_devices": n_dev, "bit_exact": bool(ok),
               "batch_sharded": {"call": "goofy_b200_encode_batch_sharded(GOOFY_B200_BOTH)", "textures": len(keep), "texture": [w, h],
                                 "value": len(keep) * w * h / dt / 1e6, "unit": "MP/s (host wall clock per call, each pixel to both codecs)"},
               "sharded_host": {"call": "goofy_b200_encode_sharded_host(DXT1)", "image": [hw, hh], "value": hw * hh / dth / 1e6,
                                "unit": "MP/s (host wall clock per call, pinned host buffers in and out)"}}
        del keep
        torch.cuda.set_device(ctx.local_rank)
    ctx.host_barrier()
    return out


def pcie_probe(ctx: Ctx, size: int):
    """What the bare copies of one end-to-end step take on this box at this N: 4 B/px pinned host -> device and 0.5 B/px
    device -> pinned host, issued concurrently on two streams, every rank at the same time, MAX over ranks."""
    torch = ctx.torch
    n_in, n_out = size * size * 4, size * size // 2
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n_in, dtype=torch.uint8, device=ctx.dev)
    d_out = torch.empty(n_out, dtype=torch.uint8, device=ctx.dev)
    s2 = torch.cuda.Stream(device=ctx.dev)

    def both():
        d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s2)
    ms, _ = ctx.timed(both, 5, 2)
    return ms / 5


def run_b200_arm(args):
    ctx = Ctx(args)
    torch, gb = ctx.torch, ctx.gb
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    lib, st = ctx.lib, ctx.stream
    timed = ctx.timed

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
    p_src, p_dst = int(src.data_ptr()), int(dst.data_ptr())

    def step(c=codec):
        # one pass over the batch through the batched device-resident entry point (one launch)
        gb.check(lib.goofy_b200_encode_batch_uniform_device(c, p_dst, p_src, size, size, stride, img_bytes, out_bytes, batch, st))

    def step_per_texture(c=codec):
        # the same batch as one call (and one launch) per 8192^2 texture
        for b in range(batch):
            gb.check(lib.goofy_b200_encode_device(c, p_dst + b * out_bytes, p_src + b * img_bytes, size, size, stride, st))

    # Clocks: one NVML reading takes milliseconds on this driver and the headline region of a 20-step run lasts 3.4 ms,
    # so the sampler (a child process) keeps going through every timed leg of this run (device-resident side legs,
    # end-to-end legs, named configs) and reports how many of its samples fell inside the headline region.
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.begin("all_legs")
    ms, launches = timed(step, args.steps, args.warmup, sampler)
    headline_kernel = gb.last_launch_kernel()

    total_px = px_per_step * args.steps * world
    value = total_px / (ms * 1e-3) / 1e6
    launches_per_step = launches / args.steps
    launch_ms = ms / launches
    bytes_per_launch = px_per_step * BYTES_PER_PIXEL / launches_per_step
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    peak, peak_src = hbm_peak()
    traffic, traffic_src = None, None   # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed ncu capture
    try:
        tj = json.loads((ROOT / "profiles" / "traffic.json").read_text())
        if size == 8192 and batch == 4:
            traffic = tj["dram_bytes_per_launch"].get(args.codec)
            traffic_src = "profiles/traffic.json: ncu --set full capture of this launch shape (" + str(tj.get("source", "committed under profiles/")) + "), not re-measured by this run"
    except Exception:
        traffic = None

    # same batch, one launch per texture (what a caller encoding single 8192^2 textures sees)
    side_steps = max(min(args.steps // 2, 100), 3)
    ms_s, launches_s = timed(step_per_texture, side_steps, 3)
    value_s = px_per_step * side_steps * world / (ms_s * 1e-3) / 1e6
    per_texture_kernel = gb.last_launch_kernel()

    # the other codec, same protocol (BASELINE.json's metric names both)
    other = gb.ETC1 if codec == gb.DXT1 else gb.DXT1
    ms_o, _ = timed(lambda: step(other), side_steps, 3)
    value_o = px_per_step * side_steps * world / (ms_o * 1e-3) / 1e6
    achieved_o = value_o / world * 1e6 * BYTES_PER_PIXEL / 1e9
    other_kernel = gb.last_launch_kernel()

    # dual-output pass: both codecs from one read (5 B/px)
    dst2 = torch.empty((batch, out_bytes), dtype=torch.uint8, device=dev)
    p_dst2 = int(dst2.data_ptr())

    def dual_step():
        gb.check(lib.goofy_b200_encode_dual_device(p_dst, p_dst2, p_src, size, size, stride, img_bytes, out_bytes, batch, st))
    ms_d, _ = timed(dual_step, side_steps, 3)
    value_d = px_per_step * side_steps * world / (ms_d * 1e-3) / 1e6
    dual_kernel = gb.last_launch_kernel()

    # packed-RGB input (3 B/px, the data format a decoded PNG / JPEG has): same batch, same protocol, 3.5 B/px
    rgb = torch.empty((batch, size, size, 3), dtype=torch.uint8, device=dev)
    rgb.copy_(src[..., :3])
    p_rgb = int(rgb.data_ptr())

    def rgb24_step():
        gb.check(lib.goofy_b200_encode_rgb24_device(codec, p_dst2, 0, p_rgb, size, size, size * 3, size * size * 3, out_bytes, batch, st))
    ms_r, _ = timed(rgb24_step, side_steps, 3)
    value_r = px_per_step * side_steps * world / (ms_r * 1e-3) / 1e6
    rgb24_kernel = gb.last_launch_kernel()
    step()
    torch.cuda.synchronize()
    rgb24_same = bool(torch.equal(dst, dst2))
    del rgb

    # what the end-to-end legs (run last, below) need of the device-resident state: one texture and its blocks
    keep_src = keep_want = None
    if not args.no_e2e:
        gb.check(gb.encode_device(codec, dst[0], src[0], size, size, stride))
        torch.cuda.synchronize()
        keep_src, keep_want = src[0].cpu(), dst[0].cpu()

    # ---- the named multi-GPU shapes (outside the headline region): BASELINE.json configs[3] and [4], and the
    #      in-library shard scheduler; every one with a bit-exactness sample against the reference
    configs = None
    sharded = None
    if not args.no_configs:
        del src, dst, dst2
        torch.cuda.empty_cache()
        cfg_steps = max(3, min(args.steps, 20))
        # (a strip launch at 8 GPUs lasts 22 us: ten times the steps, so that the timed region is milliseconds, not 0.4 ms)
        configs = {"batch1024": config_batch1024(ctx, cfg_steps, 3, args.images),
                   "strip16384": config_strip16384(ctx, cfg_steps * 10, 5)}
        sharded = sharded_api_leg(ctx)

    clocks = None
    if rank == 0:
        sampler.end("all_legs")
        clocks = sampler.finish("all_legs")

    # ---- end to end through the drop-in host API: pinned host buffers, H2D + D2H inside the timed region.  Run after
    #      the clock sampler has stopped: its child process queries NVML back to back, and every query holds a driver lock
    #      that the host path's own driver calls (copies, launches, event queries: dozens per call) then wait for (the
    #      call averaged 5.07 ms with the sampler running and 4.66-4.76 ms after it had stopped, against 4.36-4.45 ms from
    #      a C++ caller on the same boxes; round 2, sessions T and U)
    e2e = None
    if not args.no_e2e:
        h_src = torch.empty((size, size, 4), dtype=torch.uint8).pin_memory()
        h_dst = torch.empty((out_bytes,), dtype=torch.uint8).pin_memory()
        h_src.copy_(keep_src)
        host_fn = gb.compressDXT1 if codec == gb.DXT1 else gb.compressETC1

        def e2e_step():
            gb.check(host_fn(h_dst, h_src, size, size, stride))
        e2e_steps = max(3, min(args.steps, 16))
        # The same call with alpha-stripped staging turned off first (every pixel crosses the link as RGBA: the plain DMA
        # pipeline of round 1).  It doubles as the warm-up of this part of the run: the GPU and the link have idled while
        # the clock samples were parsed, and the first timed leg here measured 5 % slower than the same code run second.
        mode = gb.set_host_rgb_staging(gb.HOST_RGB_OFF)
        for _ in range(6):
            e2e_step()
        ms_raw, _ = timed(e2e_step, e2e_steps, 3)
        gb.set_host_rgb_staging(mode)
        # the host path's measured choices (pack-time estimate, packing vs plain DMA) need a few calls to settle
        for _ in range(12):
            e2e_step()
        link0 = gb.host_link_stats()
        ms_e, _ = timed(e2e_step, e2e_steps, 3)
        link1 = gb.host_link_stats()
        calls = e2e_steps + 3   # timed() ran 3 warm-up steps after link0 was read
        h2d_actual = (link1["bytes_uploaded"] - link0["bytes_uploaded"]) // calls
        pcie_ms = pcie_probe(ctx, size)
        e2e = {"value": size * size * e2e_steps * world / (ms_e * 1e-3) / 1e6, "unit": "MP/s",
               # bytes that actually crossed the link per call, counted by the library (the input tensor holds 4 B/px; the
               # host path drops the alpha byte of the strips its staging threads get to before the copy engine does)
               "h2d_bytes_per_step": h2d_actual,
               "h2d_bytes_logical": size * size * 4, "d2h_bytes_per_step": out_bytes,
               "ms_per_step": ms_e / e2e_steps,
               "host_rgb_staging": {0: "off", 1: "auto", 2: "always", 3: "pageable input only (one process per GPU on a shared host)"}[mode],
               "alpha_stripped_share_of_pixels": (size * size * 4 - h2d_actual) / (size * size),   # a stripped pixel saves one byte
               "calls_with_packing": link1["packing_calls"] - link0["packing_calls"], "calls_plain_dma": link1["plain_calls"] - link0["plain_calls"],
               "host_threads": gb.host_threads(), "host_neighbours": gb.host_neighbours(),
               "rgba_dma_only": {"value": size * size * e2e_steps * world / (ms_raw * 1e-3) / 1e6, "unit": "MP/s",
                                 "ms_per_step": ms_raw / e2e_steps, "h2d_bytes_per_step": size * size * 4,
                                 "note": "same call, goofy_b200_set_host_rgb_staging(OFF): the plain strip pipeline of round 1"},
               "pcie_bound_ms": pcie_ms, "ms_per_step_over_pcie_bound": (ms_e / e2e_steps) / pcie_ms,
               "pcie_bound_note": "bare pinned cudaMemcpyAsync of the tensors' bytes (4 B/px in, 0.5 B/px out, concurrently) on this box, "
                                  f"all {world} rank(s) at the same time, max over ranks; the host call can finish sooner because it "
                                  "sends h2d_bytes_per_step, not h2d_bytes_logical",
               "api": f"goofy_b200.compress{args.codec.upper()}(result, input, w, h, stride) on pinned host buffers"}
        # both codecs from one upload (goofy_b200_encode_dual_host): 4 B/px in, 1 B/px out
        h_dual = torch.empty((2, out_bytes), dtype=torch.uint8).pin_memory()

        def e2e_dual_step():
            gb.check(gb.encode_dual_host(h_dual[0], h_dual[1], h_src, size, size, stride))
        ms_d2, _ = timed(e2e_dual_step, e2e_steps, 2)
        e2e["dual_output_host_call"] = {"value": size * size * e2e_steps * world / (ms_d2 * 1e-3) / 1e6,
                                        "unit": "MP/s (each pixel to DXT1 AND ETC1s, one upload)", "ms_per_step": ms_d2 / e2e_steps,
                                        "d2h_bytes_per_step": 2 * out_bytes}
        # a caller whose pixels are packed RGB8 to begin with (goofy_b200_encode_rgb24_host): 3 B/px cross the link, no host work
        h_rgb = torch.empty((size, size, 3), dtype=torch.uint8).pin_memory()
        h_rgb.copy_(keep_src[..., :3])
        h_dst3 = torch.empty((out_bytes,), dtype=torch.uint8).pin_memory()

        def e2e_rgb_step():
            gb.check(gb.encode_rgb24_host(codec, h_dst3, h_rgb, size, size, size * 3))
        ms_r3, _ = timed(e2e_rgb_step, e2e_steps, 3)
        e2e["rgb24_host_call"] = {"value": size * size * e2e_steps * world / (ms_r3 * 1e-3) / 1e6, "unit": "MP/s", "ms_per_step": ms_r3 / e2e_steps,
                                  "h2d_bytes_per_step": size * size * 3, "same_bytes": bool(torch.equal(h_dst3, keep_want)),
                                  "note": "the same texture held as packed RGB8 on the host (not BASELINE.json's RGBA8 workload)"}
        del h_rgb, h_dst3
        # same call with ordinary (pageable) numpy buffers: the library stages them through pinned strips
        p_srcbuf = keep_src.numpy().reshape(-1)
        p_dstbuf = np.zeros(out_bytes, dtype=np.uint8)

        def e2e_pageable_step():
            gb.check(host_fn(p_dstbuf, p_srcbuf, size, size, stride))
        ms_p, _ = timed(e2e_pageable_step, e2e_steps, 2)
        e2e["pageable_buffers"] = {"value": size * size * e2e_steps * world / (ms_p * 1e-3) / 1e6, "unit": "MP/s",
                                   "ms_per_step": ms_p / e2e_steps}
        # the result must be the same bytes the device-resident path produced
        e2e["matches_device_path"] = bool(torch.equal(keep_want, h_dst))
        del h_src, h_dst, h_dual

    # ---- roofline.traffic, re-measured (rank 0, N=1 only; after everything that is timed)
    if rank == 0 and world == 1 and not args.no_traffic_probe and size == 8192 and batch == 4:
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        measured, why = measure_traffic_with_ncu(args.codec, size, batch)
        if measured is not None:
            traffic, traffic_src = measured, why
        elif traffic_src is not None:
            traffic_src += f" ({why})"

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
            img = np.zeros((1024, size, 4), dtype=np.uint8)
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
                       "cpu_binding": ctx.numa},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                         "kernel": headline_kernel, "kernel_source": "goofy_b200_last_launch_kernel() after the timed region",
                         "bytes_per_launch": bytes_per_launch, "launch_ms": launch_ms,
                         "like_for_like_ceiling": "tools/membench (profiles/r01_membench.txt): a trivial-compute kernel with the same "
                                                  "8:1 read:write access pattern reaches 7115 GB/s at 1.2 GB per launch "
                                                  "(6696 GB/s at 302 MB); read-only 7355 GB/s"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "per_texture_launch": {"value": value_s, "unit": "MP/s", "launches_per_step": launches_s / side_steps,
                                   "achieved_gbs_per_gpu": value_s / world * 1e6 * BYTES_PER_PIXEL / 1e9, "kernel": per_texture_kernel,
                                   "note": "same batch, one goofy_b200_encode_device call per texture"},
            "other_codec": {"codec": "etc1" if codec == gb.DXT1 else "dxt1", "value": value_o, "unit": "MP/s",
                            "achieved_gbs_per_gpu": achieved_o, "frac": achieved_o / peak, "kernel": other_kernel},
            "dual_output": {"value": value_d, "unit": "MP/s (each pixel encoded to both DXT1 and ETC1s)",
                            "achieved_gbs_per_gpu": value_d / world * 1e6 * 5.0 / 1e9, "kernel": dual_kernel},
            "rgb24_input": {"value": value_r, "unit": "MP/s", "bytes_per_pixel": 3.5,
                            "achieved_gbs_per_gpu": value_r / world * 1e6 * 3.5 / 1e9, "kernel": rgb24_kernel,
                            "same_bytes_as_rgba_path": rgb24_same,
                            "note": "goofy_b200_encode_rgb24_device: the same textures as packed RGB8 (no alpha byte), not BASELINE.json's RGBA8 workload"},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if configs is not None:
            line["configs"] = configs
            line["sharded_api"] = sharded
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if ctx.distributed:
        ctx.dist.destroy_process_group()
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
