#!/usr/bin/env python
"""A/B of several builds of libgoofy_b200.so inside ONE process (one box, one set of textures, alternated).

    python tools/ab_libs.py [--steps 200] [--rounds 3] [--check] name=path/to/lib.so [name=...]

Every library is loaded with ctypes under its own path (separate static state), and the three batched kernels
bench.py times (DXT1, ETC1s, dual-output over 4 device-resident 8192^2 textures) are timed with CUDA events on
torch's current stream.  --check compares every build's output bytes with the first build's.
Experiments only; the product loads goofy_b200/libgoofy_b200.so.
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch

from bench import fill_texture_device
from goofy_b200._lib import PROTOTYPES

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--check", action="store_true")
ap.add_argument("--out", default=None)
ap.add_argument("libs", nargs="+")
args = ap.parse_args()


def load(path):
    lib = C.CDLL(str(Path(path).resolve()))
    for name, (res, at) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, at
    return lib


libs = []
for spec in args.libs:
    name, _, path = spec.partition("=")
    # name=path@ROWS@LOADPATH: GOOFY_B200_ROWS_PER_CTA and goofy_b200_set_load_path for this build (needs its own .so file)
    path, _, rest = (path or name).partition("@")
    rows, _, lp = rest.partition("@")
    lib = load(path)
    if lp:
        lib.goofy_b200_set_load_path(int(lp))
    libs.append((name, lib, rows))

size, batch = args.size, args.batch
src = torch.empty((batch, size, size, 4), dtype=torch.uint8, device="cuda")
for b in range(batch):
    fill_texture_device(torch, src[b], seed=1 + b)
ob = size * size // 2
d1 = torch.empty((batch, ob), dtype=torch.uint8, device="cuda")
d2 = torch.empty((batch, ob), dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
px = size * size * batch


def run(lib, mode):
    if mode == "dual":
        rc = lib.goofy_b200_encode_dual_device(d1.data_ptr(), d2.data_ptr(), src.data_ptr(), size, size, size * 4, size * size * 4, ob, batch, stream)
    else:
        rc = lib.goofy_b200_encode_batch_uniform_device(0 if mode == "dxt1" else 1, d1.data_ptr(), src.data_ptr(), size, size, size * 4,
                                                        size * size * 4, ob, batch, stream)
    assert rc == 0, rc


def timed(lib, mode, steps):
    for _ in range(5):
        run(lib, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run(lib, mode)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return px * (5.0 if mode == "dual" else 4.5) / (ms * 1e-3) / 1e9


import os
for name, lib, rows in libs:   # the row-walking launcher reads its environment override on first use
    os.environ["GOOFY_B200_ROWS_PER_CTA"] = rows or "0"
    for mode in ("dxt1", "etc1", "dual"):
        run(lib, mode)
    torch.cuda.synchronize()
libs = [(name, lib) for name, lib, _ in libs]

if args.check:
    want = {}
    for name, lib in libs:
        for mode in ("dxt1", "etc1", "dual"):
            d1.zero_(); d2.zero_()
            run(lib, mode)
            torch.cuda.synchronize()
            got = (d1.clone(), d2.clone() if mode == "dual" else None)
            if mode not in want:
                want[mode] = got
            else:
                ok = torch.equal(got[0], want[mode][0]) and (got[1] is None or torch.equal(got[1], want[mode][1]))
                print(f"check {name} {mode}: {'same bytes' if ok else 'DIFFERENT'}")
                assert ok

res = {name: {m: [] for m in ("dxt1", "etc1", "dual")} for name, _ in libs}
for r in range(args.rounds):
    # rotate the order every round: under a power cap the first build timed after a pause sees higher clocks
    order = libs[r % len(libs):] + libs[:r % len(libs)]
    for mode in ("dxt1", "etc1", "dual"):
        for name, lib in order:
            res[name][mode].append(timed(lib, mode, args.steps))
for name, _ in libs:
    print(f"{name:>12}: " + " | ".join(f"{m} " + "/".join(f"{v:.0f}" for v in res[name][m]) + f" (mean {sum(res[name][m]) / len(res[name][m]):.0f})"
                                        for m in ("dxt1", "etc1", "dual")) + "  GB/s")
if args.out:
    Path(args.out).write_text(json.dumps(res, indent=1))
