import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference (oracle/_ref/libgoofy_ref.so); tests that need it skip when it was never built."""
    from oracle.oracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libgoofy_ref.so not built (needs /root/reference at build time)")
    return Reference()


@pytest.fixture(scope="session")
def kernel_math():
    """The CUDA block codec compiled for the host with emulated intrinsics (tests/kernel_math_host.cpp)."""
    import ctypes as C
    build = ROOT / "tests" / "_build"
    build.mkdir(exist_ok=True)
    so = build / "libkernel_math_host.so"
    srcs = [ROOT / "tests" / "kernel_math_host.cpp", ROOT / "goofy_b200" / "csrc" / "block_codec.cuh",
            ROOT / "goofy_b200" / "csrc" / "block_decode.cuh", ROOT / "goofy_b200" / "csrc" / "lanes.cuh"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-o", str(so), str(srcs[0])],
                       check=True, capture_output=True)
    lib = C.CDLL(str(so))
    u8p = C.POINTER(C.c_uint8)
    lib.kernel_math_compress.argtypes = [C.c_int, u8p, u8p, C.c_uint, C.c_uint, C.c_uint]
    lib.kernel_math_compress.restype = C.c_int

    import numpy as np

    def run(codec, img, width, height, stride=None):
        stride = width * 4 if stride is None else stride
        img = np.ascontiguousarray(img, dtype=np.uint8).reshape(-1)
        out = np.zeros(width * height // 2, dtype=np.uint8)
        rc = lib.kernel_math_compress(codec, out.ctypes.data_as(u8p), img.ctypes.data_as(u8p), width, height, stride)
        return rc, out

    lib.kernel_math_decode.argtypes = [C.c_int, u8p, C.c_uint, C.c_uint, u8p]
    lib.kernel_math_decode.restype = C.c_int
    lib.kernel_math_sse.argtypes = [C.c_int, u8p, u8p, C.c_uint, C.c_uint, C.POINTER(C.c_ulonglong)]
    lib.kernel_math_sse.restype = C.c_int

    def decode(codec, blocks, width, height):
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
        out = np.zeros((height, width, 4), dtype=np.uint8)
        assert lib.kernel_math_decode(codec, blocks.ctypes.data_as(u8p), width, height, out.ctypes.data_as(u8p)) == 0
        return out

    def sse(codec, blocks, img, width, height):
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
        img = np.ascontiguousarray(img, dtype=np.uint8).reshape(-1)
        acc = (C.c_ulonglong * 3)()
        assert lib.kernel_math_sse(codec, blocks.ctypes.data_as(u8p), img.ctypes.data_as(u8p), width, height, acc) == 0
        return np.array([acc[0], acc[1], acc[2]], dtype=np.float64)

    run.decode = decode
    run.sse = sse
    return run


@pytest.fixture(scope="session")
def golden():
    import json
    return json.loads((ROOT / "tests" / "golden" / "golden.json").read_text())
