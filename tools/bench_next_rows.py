#!/usr/bin/env python
"""Throughput of the SURVEY.md 8(f) rows on one B200, under bench.py's protocol (inputs resident in HBM and larger
than L2, CUDA events, >= 3 warm-up passes): float-reference flavour (N1), BC1 / ETC1 decoders and the fused
squared-error reduction (N2), relaxed shapes and ragged batches (N4).  One JSON object on stdout.

    python tools/bench_next_rows.py [--size 8192] [--steps 50]

Algorithmic bytes: encoders 4.5 B/px (4 read + 0.5 written); decoder 4.5 B/px (0.5 read + 4 written);
block SSE 4.5 B/px (0.5 + 4 read, 24 bytes written in total).
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch

import goofy_b200 as gb
from bench import fill_texture_device, hbm_peak

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--textures", type=int, default=4, help="distinct textures rotated through (4 x 256 MiB > L2)")
args = ap.parse_args()
size, n_tex = args.size, args.textures
px = size * size
ob = px // 2

src = torch.empty((n_tex, size, size, 4), dtype=torch.uint8, device="cuda")
for i in range(n_tex):
    fill_texture_device(torch, src[i], seed=7 + i)
blk = {c: torch.empty((n_tex, ob), dtype=torch.uint8, device="cuda") for c in (gb.DXT1, gb.ETC1)}
for c in blk:
    gb.check(gb.encode_batch_uniform_device(c, blk[c], src, size, size, size * 4, px * 4, ob, n_tex))
dec = torch.empty((n_tex, size, size, 4), dtype=torch.uint8, device="cuda")
out = torch.empty((n_tex, ob), dtype=torch.uint8, device="cuda")
sse = torch.zeros(3, dtype=torch.int64, device="cuda")
mips = torch.empty((n_tex, ob * 2), dtype=torch.uint8, device="cuda")   # a full chain is < 4/3 of the base level
torch.cuda.synchronize()


def timed(fn):
    for i in range(4):
        fn(i % n_tex)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        fn(i % n_tex)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps


peak, peak_src = hbm_peak()
res = {}


def record(name, ms, bytes_per_px=4.5, pixels=px):
    gbs = pixels * bytes_per_px / (ms * 1e-3) / 1e9
    res[name] = {"mp_per_s": pixels / (ms * 1e-3) / 1e6, "gb_per_s": gbs, "frac_of_measured_peak": gbs / peak, "us_per_launch": ms * 1e3}


for name, c in (("encode_dxt1", gb.DXT1), ("encode_etc1", gb.ETC1), ("encode_dxt1_floatref", gb.DXT1_FLOATREF), ("encode_etc1_floatref", gb.ETC1_FLOATREF)):
    record(name, timed(lambda i, c=c: gb.check(gb.encode_device(c, out[i], src[i], size, size, size * 4))))
for name, c in (("decode_dxt1", gb.DXT1), ("decode_etc1", gb.ETC1)):
    record(name, timed(lambda i, c=c: gb.check(gb.decode_device(c, dec[i], blk[c][i], size, size, size * 4))))
for name, c in (("block_sse_dxt1", gb.DXT1), ("block_sse_etc1", gb.ETC1)):
    record(name, timed(lambda i, c=c: gb.check(gb.block_sse_device(c, blk[c][i], src[i], size, size, size * 4, sse))))
# packed RGB input (3 bytes per pixel): 3.5 B/px of traffic; one launch per texture like the rows above, and all
# textures in one batched launch like bench.py's headline
rgb = torch.empty((n_tex, size, size, 3), dtype=torch.uint8, device="cuda")
rgb.copy_(src[..., :3])
out2 = torch.empty((n_tex, ob), dtype=torch.uint8, device="cuda")
for name, c in (("encode_rgb24_dxt1", gb.DXT1), ("encode_rgb24_etc1", gb.ETC1), ("encode_rgb24_both", gb.BOTH)):
    bpp = 4.0 if c == gb.BOTH else 3.5
    record(name, timed(lambda i, c=c: gb.check(gb.encode_rgb24_device(c, out[i], rgb[i], size, size, size * 3, d_result2=out2[i]))), bytes_per_px=bpp)
    record(name + "_batched", timed(lambda i, c=c: gb.check(gb.encode_rgb24_device(c, out, rgb, size, size, size * 3, d_result2=out2,
                                                                                   input_image_pitch=px * 3, result_image_pitch=ob, n_images=n_tex))),
           bytes_per_px=bpp, pixels=px * n_tex)
assert torch.equal(out, blk[gb.DXT1]) and torch.equal(out2, blk[gb.ETC1]), "rgb24 results differ from the RGBA path"
# relaxed shapes: the same texture minus 3 pixels in both directions (edge blocks replicate)
w2 = h2 = size - 3
for name, c in (("encode_relaxed_dxt1", gb.DXT1), ("encode_relaxed_etc1", gb.ETC1)):
    record(name, timed(lambda i, c=c: gb.check(gb.encode_relaxed_device(c, out[i], src[i], w2, h2, size * 4))), pixels=w2 * h2)
# ragged batch: a full mip chain 8192 .. 16 of every texture in one launch per texture (sources are sub-rectangles of the texture)
for name, c in (("encode_ragged_mips_dxt1", gb.DXT1), ("encode_ragged_mips_etc1", gb.ETC1)):
    chains = []
    total_px = 0
    for i in range(n_tex):
        items, off, s = [], 0, size
        while s >= 16:
            items.append((src[i].data_ptr(), mips[i].data_ptr() + off, s, s, size * 4))
            off += s * s // 2
            if i == 0:
                total_px += s * s
            s //= 2
        chains.append(gb.make_descriptors(items))
    record(name, timed(lambda i, c=c: gb.check(gb.encode_batch_device(c, chains[i]))), pixels=total_px)

print(json.dumps({"workload": f"{size}x{size} RGBA8, {n_tex} textures rotated, one launch per texture, {args.steps} launches per figure",
                  "peak_gb_per_s": peak, "peak_source": peak_src, "results": res}))
