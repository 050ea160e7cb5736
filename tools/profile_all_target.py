#!/usr/bin/env python
"""Launch sequence for one `ncu --set full` pass over every kernel that is NOT one of the AUTO encoders of the headline
(those are covered by tools/profile_target.py): TMA tile layer (3 modes, on the padded-stride strip it is meant for),
float-reference flavours, relaxed shapes, ragged batch (mip chain, both codecs in one launch), pitched row-walking batch,
short-launch instantiations, BC1 / ETC1 decoders, fused block SSE.  Every kernel is launched twice (the second launch
is the one to read: -s / -c of the ncu command line skip the first)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

import goofy_b200 as gb
from bench import fill_texture_device

size = 8192
px = size * size
src = torch.empty((size, size, 4), dtype=torch.uint8, device="cuda")
fill_texture_device(torch, src, seed=3)
out = torch.empty(px // 2, dtype=torch.uint8, device="cuda")
out2 = torch.empty(px // 2, dtype=torch.uint8, device="cuda")
dec = torch.empty((size, size, 4), dtype=torch.uint8, device="cuda")
sse = torch.zeros(3, dtype=torch.int64, device="cuda")
# the strip of BASELINE.json configs[4] at 8 GPUs: 16384 x 2048, stride 65792, pad 0xAB
sw, sh, sstride = 16384, 2048, 16384 * 4 + 256
strip = torch.full((sh, sstride), 0xAB, dtype=torch.uint8, device="cuda")
strip[:, : sw * 4] = src.view(-1)[: sh * sw * 4].view(sh, sw * 4)
sout = torch.empty(sw * sh // 2, dtype=torch.uint8, device="cuda")
sout2 = torch.empty(sw * sh // 2, dtype=torch.uint8, device="cuda")
# pitched batch: 64 x 1024^2, images 4 KiB apart
bw_ = bh_ = 1024
pitch = bw_ * bh_ * 4 + 4096
batch = torch.full((64 * pitch,), 0xAB, dtype=torch.uint8, device="cuda")
for i in range(64):
    batch[i * pitch: i * pitch + bw_ * bh_ * 4].copy_(src.view(-1)[i * bw_ * bh_ * 4: (i + 1) * bw_ * bh_ * 4])
bout = torch.empty(64 * bw_ * bh_ // 2, dtype=torch.uint8, device="cuda")
bout2 = torch.empty(64 * bw_ * bh_ // 2, dtype=torch.uint8, device="cuda")
mips = torch.empty(px, dtype=torch.uint8, device="cuda")
mips2 = torch.empty(px, dtype=torch.uint8, device="cuda")
chain, off, s = [], 0, size
while s >= 16:
    chain.append((src.data_ptr(), mips.data_ptr() + off, s, s, size * 4, -1, mips2.data_ptr() + off))
    off += s * s // 2
    s //= 2
chain = gb.make_descriptors(chain)
torch.cuda.synchronize()


def twice(fn):
    for _ in range(2):
        gb.check(fn())
    torch.cuda.synchronize()


prev = gb.set_load_path(gb.LOAD_TMA)
for codec in (gb.DXT1, gb.ETC1):
    twice(lambda: gb.encode_device(codec, sout, strip, sw, sh, sstride))
twice(lambda: gb.encode_dual_device(sout, sout2, strip, sw, sh, sstride))
gb.set_load_path(prev)
# AUTO on the same strip: the short-launch instantiations
for codec in (gb.DXT1, gb.ETC1):
    twice(lambda: gb.encode_device(codec, sout, strip, sw, sh, sstride))
twice(lambda: gb.encode_dual_device(sout, sout2, strip, sw, sh, sstride))
# pitched batch: row-walking kernels with grid.z = image
twice(lambda: gb.encode_batch_uniform_device(gb.ETC1, bout, batch, bw_, bh_, bw_ * 4, pitch, bw_ * bh_ // 2, 64))
twice(lambda: gb.encode_dual_device(bout, bout2, batch, bw_, bh_, bw_ * 4, pitch, bw_ * bh_ // 2, 64))
for codec in (gb.DXT1_FLOATREF, gb.ETC1_FLOATREF):
    twice(lambda: gb.encode_device(codec, out, src, size, size, size * 4))
for codec in (gb.DXT1, gb.ETC1):
    twice(lambda: gb.encode_relaxed_device(codec, out, src, size - 3, size - 3, size * 4))
for codec in (gb.DXT1, gb.ETC1, gb.BOTH):
    twice(lambda: gb.encode_batch_device(codec, chain))
for codec in (gb.DXT1, gb.ETC1):
    gb.check(gb.encode_device(codec, out, src, size, size, size * 4))
    twice(lambda: gb.decode_device(codec, dec, out, size, size, size * 4))
    twice(lambda: gb.block_sse_device(codec, out, src, size, size, size * 4, sse))
print("done")
