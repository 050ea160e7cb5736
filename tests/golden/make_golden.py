#!/usr/bin/env python
"""Regenerates tests/golden/ from the UNMODIFIED reference (oracle/_ref/libgoofy_ref.so, built by
oracle/Makefile from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs
  golden.json    sha256 / crc32 of goofy::compressDXT1/ETC1 output for every loadable test image
                 (tight stride, alpha 255) and the seeded synthetic textures, + known-answer bytes
  fixtures.npz   small inputs with their full expected outputs, so the oracle stays pinned on
                 machines that have neither the reference nor its test images
"""
import hashlib
import json
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.oracle import (DXT1, ETC1, Reference, aligned_copy, load_test_image, synth_family,  # noqa: E402
                           image_names, xorshift_bytes)

HERE = Path(__file__).resolve().parent


def digest(b: np.ndarray) -> dict:
    raw = b.tobytes()
    return {"sha256": hashlib.sha256(raw).hexdigest(), "crc32": f"{zlib.crc32(raw) & 0xFFFFFFFF:08x}", "bytes": len(raw)}


def main():
    ref = Reference()
    out = {"generator": "tests/golden/make_golden.py", "source": "goofy::compressDXT1/ETC1, GoofyTC/goofy_tc.h:1497-1557, g++ -O2 -msse2",
           "images": {}, "synthetic": {}, "known_answer": {}}
    fixtures = {}

    for name in image_names():
        img = load_test_image(name)
        h, w = img.shape[:2]
        flat = aligned_copy(img)
        entry = {"width": w, "height": h, "input_sha256": hashlib.sha256(img.tobytes()).hexdigest()}
        for codec, key in ((DXT1, "dxt1"), (ETC1, "etc1")):
            rc, blocks = ref.compress(codec, flat, w, h)
            assert rc == 0
            entry[key] = digest(blocks)
        out["images"][name] = entry
        # 64x32 crop from the middle of every image as a self-contained fixture (patterns: whole image)
        ch, cw = (h, w) if name == "patterns" else (32, 64)
        y0, x0 = ((h - ch) // 2) & ~3, ((w - cw) // 2) & ~15
        crop = np.ascontiguousarray(img[y0:y0 + ch, x0:x0 + cw])
        fixtures[f"img_{name}_rgba"] = crop
        for codec, key in ((DXT1, "dxt1"), (ETC1, "etc1")):
            rc, blocks = ref.compress(codec, aligned_copy(crop), cw, ch)
            assert rc == 0
            fixtures[f"img_{name}_{key}"] = blocks

    # SURVEY.md Appendix B synthetic textures (sequential xorshift64* stream)
    W = H = 1024
    for k in range(3):
        v = xorshift_bytes(W * H * 4, k).astype(np.int64)
        i = np.arange(W * H * 4, dtype=np.int64)
        x, y = (i // 4) % W, (i // 4) // W
        if k == 0:
            data = v
        elif k == 1:
            data = ((x + y) // 8 + (v & 15) + (i & 3) * 20) & 255
        else:
            data = np.where(v & 1, 255, 0)
        data = data.astype(np.uint8)
        entry = {"width": W, "height": H, "input_sha256": hashlib.sha256(data.tobytes()).hexdigest()}
        for codec, key in ((DXT1, "dxt1"), (ETC1, "etc1")):
            rc, blocks = ref.compress(codec, aligned_copy(data), W, H)
            assert rc == 0
            entry[key] = digest(blocks)
        out["synthetic"][f"synth{k}"] = entry

    # counter-based families used by the tests (oracle.synth_family), 256x256
    for fam in range(4):
        img = synth_family(fam, 256, 256)
        entry = {"width": 256, "height": 256, "input_sha256": hashlib.sha256(img.tobytes()).hexdigest()}
        for codec, key in ((DXT1, "dxt1"), (ETC1, "etc1")):
            rc, blocks = ref.compress(codec, aligned_copy(img), 256, 256)
            assert rc == 0
            entry[key] = digest(blocks)
        out["synthetic"][f"family{fam}_256"] = entry

    # edge-case fixtures: constant blocks, extremes, low ranges around the range clamp and the ETC1 table steps
    rng = np.random.default_rng(20261017)
    edge = np.zeros((64, 256, 4), dtype=np.uint8)
    blocks_y, blocks_x = 16, 64
    for by in range(blocks_y):
        for bx in range(blocks_x):
            k = by * blocks_x + bx
            base = rng.integers(0, 256, size=3)
            spread = [0, 1, 2, 7, 8, 9, 15, 16, 21, 22, 23, 43, 44, 45, 73, 74, 105, 106, 151, 152, 181, 182, 253, 254, 255][k % 25]
            blk = base[None, None, :] + rng.integers(0, spread + 1, size=(4, 4, 3))
            if k % 7 == 0:
                blk = np.where(rng.integers(0, 2, size=(4, 4, 3)) > 0, 255, 0)
            if k % 11 == 0:
                blk = np.full((4, 4, 3), int(base[0]))
            edge[4 * by:4 * by + 4, 4 * bx:4 * bx + 4, :3] = np.clip(blk, 0, 255)
            edge[4 * by:4 * by + 4, 4 * bx:4 * bx + 4, 3] = rng.integers(0, 256, size=(4, 4))
    fixtures["edge_rgba"] = edge
    for codec, key in ((DXT1, "dxt1"), (ETC1, "etc1")):
        rc, blocks = ref.compress(codec, aligned_copy(edge), 256, 64)
        assert rc == 0
        fixtures[f"edge_{key}"] = blocks

    pat = load_test_image("patterns")
    for codec, key in ((DXT1, "dxt1"), (ETC1, "etc1")):
        rc, blocks = ref.compress(codec, aligned_copy(pat), 32, 32)
        out["known_answer"][f"patterns_{key}_first32"] = blocks[:32].tobytes().hex()

    (HERE / "golden.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    np.savez_compressed(HERE / "fixtures.npz", **fixtures)
    print("images", len(out["images"]), "fixtures", len(fixtures), "npz bytes", (HERE / "fixtures.npz").stat().st_size)


if __name__ == "__main__":
    main()
