"""The C-ABI library loads and exports exactly what include/goofy_b200.h declares; host-side
argument checks behave like the reference's without touching a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import goofy_b200 as gb
from goofy_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    hdr = (ROOT / "include" / "goofy_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(goofy_b200_\w+)\s*\(", hdr)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    assert "goofy_b200_compress_dxt1" in names and "goofy_b200_compress_etc1" in names
    assert len(names) == len(_lib.PROTOTYPES)
    assert set(names) == set(_lib.PROTOTYPES)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.goofy_b200_abi_version() == 3


def test_cpp_header_keeps_reference_signatures():
    """include/goofy_tc.h declares goofy::compressDXT1/ETC1 exactly as GoofyTC/goofy_tc.h:11-12 does."""
    txt = (ROOT / "include" / "goofy_tc.h").read_text()
    for fn in ("compressDXT1", "compressETC1"):
        assert re.search(rf"int {fn}\(unsigned char\* result, const unsigned char\* input, unsigned int width, "
                         rf"unsigned int height, unsigned int stride\)", txt), fn


def test_error_strings():
    assert gb.error_string(0) == "ok"
    assert "16" in gb.error_string(-1) and "4" in gb.error_string(-2)
    for code in (-3, -4, -5, -6, -7, -8, -101, -5000):
        assert gb.error_string(code)


def test_shape_checks_come_first_and_never_need_a_gpu():
    """Same order as goofy_tc.h:1500-1508: width, then height; empty images succeed."""
    out = np.full(64, 0x5A, dtype=np.uint8)
    img = np.zeros(64 * 64 * 4, dtype=np.uint8)
    for fn in (gb.compressDXT1, gb.compressETC1):
        assert fn(out, img, 24, 32, 96) == -1
        assert fn(out, img, 32, 30, 128) == -2
        assert fn(out, img, 24, 30, 96) == -1
        assert fn(out, img, 0, 0, 0) == 0
        assert fn(out, img, 32, 32, 64) == -5
        assert fn(out, img, 32, 32, 132) == -4
    assert gb.encode_host(9, out, img, 32, 32, 128) == -6
    assert gb.encode_device(0, 0, 0, 24, 32, 96) == -1
    assert gb.encode_batch_uniform_device(0, 0, 0, 32, 6, 128, 0, 0, 3) == -2
    assert gb.encode_dual_device(0, 0, 0, 0, 0, 0) == 0
    assert gb.encode_batch_device(0, [], None) == 0
    assert (out == 0x5A).all()


def test_no_cpu_fallback_without_a_gpu():
    if gb.device_count() > 0:
        pytest.skip("a GPU is present")
    out = np.zeros(512, dtype=np.uint8)
    img = np.zeros(32 * 32 * 4, dtype=np.uint8)
    assert gb.compressDXT1(out, img, 32, 32, 128) == -7   # GOOFY_B200_E_DEVICE, output untouched
    assert not out.any()
    assert gb.encode_sharded_host(0, out, img, 32, 32, 128, 0) == -7
    assert gb.encode_host_batch(0, [(img, out, 32, 32, 128)]) == -7
    assert gb.encode_dual_host(out, out, img, 32, 32, 128) == -7


def test_host_batch_validates_every_image_before_starting():
    out = np.zeros(512, dtype=np.uint8)
    img = np.zeros(32 * 32 * 4, dtype=np.uint8)
    assert gb.encode_host_batch(0, []) == 0
    assert gb.encode_host_batch(9, [(img, out, 32, 32, 128)]) == -6                              # codec
    assert gb.encode_host_batch(0, [(img, out, 32, 32, 128), (img, out, 24, 32, 128)]) == -1     # second image: width % 16
    assert gb.encode_host_batch(1, [(img, out, 32, 30, 128)]) == -2                              # height % 4
    assert gb.encode_host_batch(0, [(img, out, 32, 32, 64)]) == -5                               # stride < width * 4
    assert gb.encode_host_batch(0, [(None, out, 32, 32, 128)]) == -3                             # null input
    assert gb.encode_dual_host(out, out, img, 24, 32, 128) == -1 and gb.encode_dual_host(out, out, img, 32, 30, 128) == -2
    assert gb.encode_dual_host(out, None, img, 32, 32, 128) == -3 and gb.encode_dual_host(out, out, img, 0, 0, 0) == 0
    assert not out.any()


def test_strip_partition_matches_python_twin():
    from goofy_b200 import sharding
    for h in (4, 8, 36, 512, 2048, 16384):
        for n in (1, 2, 3, 4, 8):
            rows = 0
            for g in range(n):
                first, count = gb.strip_partition(h, n, g)
                assert (first, count) == sharding.strip_partition(h, n, g)
                assert first == rows
                rows += count
            assert rows == h // 4
    assert gb.strip_partition(64, 4, 7) == (0, 0)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    for p in list((ROOT / "goofy_b200").rglob("*.py")) + list((ROOT / "goofy_b200" / "csrc").glob("*")) + \
            list((ROOT / "include").glob("*")) + list((ROOT / "Src").glob("*")):
        if p.is_file() and p.suffix not in (".so", ".o"):
            txt = p.read_text(errors="ignore")
            assert "oracle" not in txt.lower() or "never" in txt.lower(), p
