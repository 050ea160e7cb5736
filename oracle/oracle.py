"""ctypes bindings for the CPU oracles -- TEST INFRASTRUCTURE ONLY.

Two checkers live here:

* ``Oracle``   -- oracle/liboracle.so, our scalar C restatement (oracle/goofy_oracle.c).
* ``Reference``-- oracle/_ref/libgoofy_ref.so, the unmodified reference
  (``goofy::compressDXT1/ETC1``, GoofyTC/goofy_tc.h:1497-1557) behind oracle/ref_shim.cpp.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``goofy_b200``) never does: it fails loudly when the CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
REF_DIR = ORACLE_DIR / "_ref"
TEST_DATA_DIR = REF_DIR / "test-data"

DXT1, ETC1 = 0, 1
CODEC_NAMES = {DXT1: "dxt1", ETC1: "etc1"}

_u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> None:
    """Run oracle/Makefile (liboracle.so always; _ref only when /root/reference exists)."""
    if force or not (ORACLE_DIR / "liboracle.so").exists() or (
        Path("/root/reference/GoofyTC").is_dir() and not (REF_DIR / "libgoofy_ref.so").exists()
    ):
        subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_u8p)


def _check_image(img: np.ndarray, width: int, height: int, stride: int) -> None:
    assert img.dtype == np.uint8 and img.flags["C_CONTIGUOUS"]
    need = (height - 1) * stride + width * 4 if height and width else 0
    assert img.size >= need, (img.size, need)


class _Encoder:
    """Shared numpy-facing surface of both checkers."""

    _fn = {}

    def compress(self, codec: int, img: np.ndarray, width: int, height: int, stride: int | None = None):
        """Returns (rc, blocks) with blocks a uint8 array of width*height/2 bytes (untouched on error)."""
        stride = width * 4 if stride is None else stride
        img = np.ascontiguousarray(img).reshape(-1)
        _check_image(img, width, height, stride)
        out = np.zeros(width * height // 2, dtype=np.uint8)
        rc = self._fn[codec](_ptr(out), _ptr(img), width, height, stride)
        return rc, out


class Oracle(_Encoder):
    def __init__(self):
        build()
        lib = C.CDLL(str(ORACLE_DIR / "liboracle.so"))
        sig = [_u8p, _u8p, C.c_uint, C.c_uint, C.c_uint]
        for name in ("goofy_oracle_compress_dxt1", "goofy_oracle_compress_etc1",
                     "goofy_oracle_floatref_compress_dxt1", "goofy_oracle_floatref_compress_etc1"):
            getattr(lib, name).argtypes = sig
            getattr(lib, name).restype = C.c_int
        for name in ("goofy_oracle_decode_dxt1", "goofy_oracle_decode_etc1"):
            getattr(lib, name).argtypes = [_u8p, C.c_uint, C.c_uint, _u8p]
            getattr(lib, name).restype = None
        lib.goofy_oracle_sse_rgb.argtypes = [_u8p, _u8p, C.c_size_t, C.POINTER(C.c_uint64)]
        lib.goofy_oracle_sse_rgb.restype = None
        self.lib = lib
        self._fn = {DXT1: lib.goofy_oracle_compress_dxt1, ETC1: lib.goofy_oracle_compress_etc1}
        self._dec = {DXT1: lib.goofy_oracle_decode_dxt1, ETC1: lib.goofy_oracle_decode_etc1}
        self._floatref = {DXT1: lib.goofy_oracle_floatref_compress_dxt1, ETC1: lib.goofy_oracle_floatref_compress_etc1}

    def compress_float_reference(self, codec: int, img: np.ndarray, width: int, height: int, stride: int | None = None):
        """Restatement of goofyRef:: (Src/goofy_tc_reference.cpp), the reference's float flavour."""
        stride = width * 4 if stride is None else stride
        img = np.ascontiguousarray(img).reshape(-1)
        _check_image(img, width, height, stride)
        out = np.zeros(width * height // 2, dtype=np.uint8)
        rc = self._floatref[codec](_ptr(out), _ptr(img), width, height, stride)
        return rc, out

    def decode(self, codec: int, blocks: np.ndarray, width: int, height: int) -> np.ndarray:
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
        assert blocks.size == width * height // 2
        out = np.zeros((height, width, 4), dtype=np.uint8)
        self._dec[codec](_ptr(blocks), width, height, _ptr(out))
        return out

    def sse_rgb(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
        b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1)
        assert a.size == b.size and a.size % 4 == 0
        sse = (C.c_uint64 * 3)()
        self.lib.goofy_oracle_sse_rgb(_ptr(a), _ptr(b), a.size // 4, sse)
        return np.array(list(sse), dtype=np.float64)


class Reference(_Encoder):
    """The unmodified reference.  ``Reference.available()`` is False where oracle/_ref was never built."""

    @staticmethod
    def available() -> bool:
        build()
        return (REF_DIR / "libgoofy_ref.so").exists()

    def __init__(self):
        if not self.available():
            raise FileNotFoundError("oracle/_ref/libgoofy_ref.so missing (needs /root/reference at build time)")
        lib = C.CDLL(str(REF_DIR / "libgoofy_ref.so"))
        sig = [_u8p, _u8p, C.c_uint, C.c_uint, C.c_uint]
        for name in ("ref_goofy_compress_dxt1", "ref_goofy_compress_etc1",
                     "ref_goofyref_compress_dxt1", "ref_goofyref_compress_etc1"):
            getattr(lib, name).argtypes = sig
            getattr(lib, name).restype = C.c_int
        for name in ("ref_decode_dxt1", "ref_decode_etc1"):
            getattr(lib, name).argtypes = [_u8p, C.c_uint, C.c_uint, _u8p]
            getattr(lib, name).restype = None
        lib.ref_goofy_compress_mt.argtypes = [C.c_int] + sig + [C.c_int]
        lib.ref_goofy_compress_mt.restype = C.c_int
        lib.ref_hardware_threads.restype = C.c_uint
        self.lib = lib
        self._fn = {DXT1: lib.ref_goofy_compress_dxt1, ETC1: lib.ref_goofy_compress_etc1}
        self._float = {DXT1: lib.ref_goofyref_compress_dxt1, ETC1: lib.ref_goofyref_compress_etc1}
        self._dec = {DXT1: lib.ref_decode_dxt1, ETC1: lib.ref_decode_etc1}

    def compress_float_reference(self, codec: int, img: np.ndarray, width: int, height: int):
        """goofyRef:: (Src/goofy_tc_reference.cpp) -- cross-check only, tight stride only."""
        img = np.ascontiguousarray(img).reshape(-1)
        out = np.zeros(width * height // 2, dtype=np.uint8)
        rc = self._float[codec](_ptr(out), _ptr(img), width, height, width * 4)
        return rc, out

    def compress_mt(self, codec: int, img: np.ndarray, width: int, height: int, stride: int, threads: int,
                    out: np.ndarray | None = None):
        img = np.ascontiguousarray(img).reshape(-1)
        if out is None:
            out = np.zeros(width * height // 2, dtype=np.uint8)
        rc = self.lib.ref_goofy_compress_mt(codec, _ptr(out), _ptr(img), width, height, stride, threads)
        return rc, out

    def decode(self, codec: int, blocks: np.ndarray, width: int, height: int) -> np.ndarray:
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
        out = np.zeros((height, width, 4), dtype=np.uint8)
        self._dec[codec](_ptr(blocks), width, height, _ptr(out))
        return out

    def hardware_threads(self) -> int:
        return int(self.lib.ref_hardware_threads())


# --------------------------------------------------------------------------- inputs

def aligned_empty(nbytes: int, align: int = 64) -> np.ndarray:
    """uint8 buffer whose address is a multiple of `align` (the reference uses aligned SSE loads)."""
    raw = np.empty(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]


def aligned_copy(a: np.ndarray, align: int = 64) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
    out = aligned_empty(a.size, align)
    out[:] = a
    return out


def xorshift_bytes(n: int, k: int) -> np.ndarray:
    """The sequential xorshift64* byte stream SURVEY.md Appendix B defines (state 0x9E3779B97F4A7C15 + k)."""
    s = (0x9E3779B97F4A7C15 + k) & 0xFFFFFFFFFFFFFFFF
    out = np.empty(n, dtype=np.uint8)
    M = 0xFFFFFFFFFFFFFFFF
    for i in range(n):
        s ^= s >> 12
        s ^= (s << 25) & M
        s ^= s >> 27
        out[i] = ((s * 2685821657736338717) & M) >> 56
    return out


def splitmix_rgba(n_pixels: int, seed: int, first_pixel: int = 0) -> np.ndarray:
    """Counter-based generator: pixel i = low 32 bits of splitmix64(seed + i*golden).
    Vectorised, so host and device can make any pixel independently.  Returns uint8 [n_pixels*4]."""
    with np.errstate(over="ignore"):
        i = np.arange(first_pixel, first_pixel + n_pixels, dtype=np.uint64)
        z = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z & np.uint64(0xFFFFFFFF)).astype("<u4").view(np.uint8)


def synth_family(family: int, width: int, height: int, seed: int = 0x9E3779B97F4A7C15) -> np.ndarray:
    """Seeded synthetic RGBA8 textures (tight stride), SURVEY.md section 8(d):
    0 uniform random, 1 smooth gradient + 4-bit noise, 2 binary 0/255, 3 low-range blocks."""
    n = width * height
    rnd = splitmix_rgba(n, seed + family).reshape(height, width, 4)
    if family == 0:
        return rnd.copy()
    yy, xx = np.mgrid[0:height, 0:width]
    if family == 1:
        base = ((xx + yy) // 8)[..., None] + (rnd & 15) + (np.arange(4) * 20)[None, None, :]
        return (base & 255).astype(np.uint8)
    if family == 2:
        return np.where(rnd & 1, 255, 0).astype(np.uint8)
    if family == 3:
        k = ((xx // 4 + yy // 4) % 24)[..., None]
        blockbase = splitmix_rgba((height // 4) * (width // 4), seed ^ 0x5555).reshape(height // 4, width // 4, 4)
        blockbase = np.repeat(np.repeat(blockbase, 4, axis=0), 4, axis=1)
        return ((blockbase.astype(np.int64) + rnd % (1 + k)) & 255).astype(np.uint8)
    raise ValueError(family)


def load_test_image(name: str) -> np.ndarray:
    """RGBA8 [h, w, 4] with alpha forced to 255 like the reference loader (Src/main.cpp:328-335)."""
    from PIL import Image

    im = Image.open(TEST_DATA_DIR / f"{name}.png").convert("RGBA")
    a = np.array(im, dtype=np.uint8)
    a[..., 3] = 255
    return a


def image_names() -> list[str]:
    """Images the reference harness can load: width%16==0 and height%4==0 (Src/main.cpp:298-307)."""
    if not TEST_DATA_DIR.is_dir():
        return []
    from PIL import Image

    names = []
    for p in sorted(TEST_DATA_DIR.glob("*.png")):
        with Image.open(p) as im:
            w, h = im.size
        if w % 16 == 0 and h % 4 == 0:
            names.append(p.stem)
    return names


def psnr_from_sse(sse_rgb: np.ndarray, pixels: int) -> dict:
    """psnr_rgb768 follows Src/main.cpp:444,466 (peak 768 over the SUM of channel MSEs)."""
    mse_sum = float(sse_rgb.sum()) / pixels
    if mse_sum == 0:
        return {"psnr_rgb768": float("inf"), "psnr_textbook": float("inf")}
    return {
        "psnr_rgb768": 10.0 * np.log10(768.0 * 768.0 / mse_sum),
        "psnr_textbook": 10.0 * np.log10(255.0 * 255.0 / (mse_sum / 3.0)),
    }
