"""include/goofy_png.h (the harness's PNG ingest) against PIL: every test image of the reference, plus synthetic PNGs
of the other colour types / bit depths / filters, and the reference loader's contract (Src/main.cpp:258-343 of the
reference: alpha := 0xFF, 64-byte alignment, width % 16 and height % 4 rejected)."""
import ctypes as C
import io
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
Image = pytest.importorskip("PIL.Image")


@pytest.fixture(scope="module")
def png_probe():
    build = ROOT / "tests" / "_build"
    build.mkdir(exist_ok=True)
    so = build / "libpng_loader_host.so"
    srcs = [ROOT / "tests" / "png_loader_host.cpp", ROOT / "include" / "goofy_png.h"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(so), str(srcs[0])], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    lib.goofy_png_probe.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.c_void_p, C.c_ulonglong, C.c_char_p, C.c_int]
    lib.goofy_png_probe.restype = C.c_int

    def load(path, require_shape=True, capacity=64 << 20):
        w, h = C.c_uint(), C.c_uint()
        out = np.zeros(capacity, dtype=np.uint8)
        err = C.create_string_buffer(256)
        rc = lib.goofy_png_probe(str(path).encode(), int(require_shape), C.byref(w), C.byref(h), out.ctypes.data, capacity, err, 256)
        if rc == 1:
            return None, err.value.decode()
        assert rc == 0, "buffer is not 64-byte aligned"
        return out[: w.value * h.value * 4].reshape(h.value, w.value, 4), ""
    return load


def pil_rgba_opaque(path_or_file):
    a = np.array(Image.open(path_or_file).convert("RGBA"), dtype=np.uint8)
    a[..., 3] = 255
    return a


def test_reference_test_images_decode_like_pil(png_probe):
    data = ROOT / "oracle" / "_ref" / "test-data"
    files = sorted(data.glob("*.png"))
    if not files:
        pytest.skip("oracle/_ref/test-data not present")
    loaded = rejected = 0
    for f in files:
        want = pil_rgba_opaque(f)
        got, err = png_probe(f)
        if want.shape[1] % 16 or want.shape[0] % 4:
            assert got is None and "multiple of" in err, f.name     # the reference loader's shape check
            got, err = png_probe(f, require_shape=False)
            rejected += 1
        assert got is not None, (f.name, err)
        assert np.array_equal(got, want), f.name
        loaded += 1
    assert loaded >= 38 and rejected >= 1


@pytest.mark.parametrize("mode", ["RGB", "RGBA", "L", "LA", "P", "1", "I;16"])
def test_other_colour_types_and_filters(png_probe, tmp_path, mode):
    rng = np.random.default_rng(7)
    w, h = 48, 20
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = ((xx * 5 + yy * 3) & 255).astype(np.uint8)      # smooth rows make the encoder pick Sub / Up / Paeth filters
    if mode == "RGB":
        im = Image.fromarray(np.stack([smooth, smooth[::-1], rng.integers(0, 256, (h, w), dtype=np.uint8)], axis=-1), "RGB")
    elif mode == "RGBA":
        im = Image.fromarray(np.stack([smooth, smooth.T[:h, :w] if False else smooth, smooth[::-1], rng.integers(0, 256, (h, w), dtype=np.uint8)], axis=-1), "RGBA")
    elif mode == "L":
        im = Image.fromarray(smooth, "L")
    elif mode == "LA":
        im = Image.fromarray(np.stack([smooth, smooth[::-1]], axis=-1), "LA")
    elif mode == "P":
        im = Image.fromarray(np.stack([smooth, smooth[::-1], smooth], axis=-1), "RGB").quantize(64)
    elif mode == "1":
        im = Image.fromarray((smooth > 100).astype(np.uint8) * 255, "L").convert("1")
    else:
        im = Image.fromarray((smooth.astype(np.uint16) << 8 | 0x17), "I;16")
    path = tmp_path / f"t_{mode.replace(';', '')}.png"
    im.save(path, optimize=(mode in ("RGB", "P")))
    got, err = png_probe(path, require_shape=False)
    assert got is not None, err
    if mode == "I;16":
        want = np.repeat(smooth[..., None], 4, axis=-1)
        want[..., 3] = 255
    else:
        want = pil_rgba_opaque(path)
    assert np.array_equal(got, want), mode


def test_garbage_is_rejected(png_probe, tmp_path):
    p = tmp_path / "x.png"
    p.write_bytes(b"not a png at all")
    assert png_probe(p)[0] is None
    buf = io.BytesIO()
    Image.fromarray(np.zeros((8, 16, 3), dtype=np.uint8), "RGB").save(buf, format="PNG")
    raw = bytearray(buf.getvalue())
    p.write_bytes(bytes(raw[: len(raw) // 2]))
    assert png_probe(p, require_shape=False)[0] is None
    assert png_probe(tmp_path / "missing.png")[0] is None
