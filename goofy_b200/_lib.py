"""ctypes binding of libgoofy_b200.so -- one prototype per declaration in include/goofy_b200.h.

No fallback of any kind: if the library is missing or cannot be loaded, importing the encoders
raises.  (The library itself returns GOOFY_B200_E_DEVICE when there is no CUDA device.)
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libgoofy_b200.so"


class GoofyB200Image(C.Structure):
    """struct GoofyB200Image (include/goofy_b200.h)."""
    _fields_ = [
        ("src", C.c_void_p),
        ("dst", C.c_void_p),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("stride", C.c_uint32),
        ("device", C.c_int32),
        ("dst2", C.c_void_p),
    ]


_vp, _u32, _u64, _int = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int

# name -> (restype, argtypes); tests/test_cabi.py checks this table against the header
PROTOTYPES = {
    "goofy_b200_abi_version": (_int, []),
    "goofy_b200_device_count": (_int, []),
    "goofy_b200_error_string": (C.c_char_p, [_int]),
    "goofy_b200_kernel_launches": (_u64, []),
    "goofy_b200_host_scratch_sets": (_u64, []),
    "goofy_b200_last_launch_kernel": (C.c_char_p, []),
    "goofy_b200_set_load_path": (_int, [_int]),
    "goofy_b200_get_load_path": (_int, []),
    "goofy_b200_set_host_rgb_staging": (_int, [_int]),
    "goofy_b200_get_host_rgb_staging": (_int, []),
    "goofy_b200_host_threads": (_int, []),
    "goofy_b200_host_neighbours": (_int, []),
    "goofy_b200_host_link_stats": (None, [C.POINTER(_u64)] * 5),
    "goofy_b200_compress_dxt1": (_int, [_vp, _vp, C.c_uint, C.c_uint, C.c_uint]),
    "goofy_b200_compress_etc1": (_int, [_vp, _vp, C.c_uint, C.c_uint, C.c_uint]),
    "goofy_b200_compress_dxt1_floatref": (_int, [_vp, _vp, C.c_uint, C.c_uint, C.c_uint]),
    "goofy_b200_compress_etc1_floatref": (_int, [_vp, _vp, C.c_uint, C.c_uint, C.c_uint]),
    "goofy_b200_encode_host": (_int, [_int, _vp, _vp, _u32, _u32, _u32]),
    "goofy_b200_encode_dual_host": (_int, [_vp, _vp, _vp, _u32, _u32, _u32]),
    "goofy_b200_encode_rgb24_host": (_int, [_int, _vp, _vp, _vp, _u32, _u32, _u32]),
    "goofy_b200_encode_host_batch": (_int, [_int, C.POINTER(GoofyB200Image), _u32]),
    "goofy_b200_encode_device": (_int, [_int, _vp, _vp, _u32, _u32, _u32, _vp]),
    "goofy_b200_encode_rgb24_device": (_int, [_int, _vp, _vp, _vp, _u32, _u32, _u32, _u64, _u64, _u32, _vp]),
    "goofy_b200_encode_relaxed_device": (_int, [_int, _vp, _vp, _u32, _u32, _u32, _vp]),
    "goofy_b200_encode_batch_uniform_device": (_int, [_int, _vp, _vp, _u32, _u32, _u32, _u64, _u64, _u32, _vp]),
    "goofy_b200_encode_dual_device": (_int, [_vp, _vp, _vp, _u32, _u32, _u32, _u64, _u64, _u32, _vp]),
    "goofy_b200_decode_device": (_int, [_int, _vp, _vp, _u32, _u32, _u32, _vp]),
    "goofy_b200_block_sse_device": (_int, [_int, _vp, _vp, _u32, _u32, _u32, _vp, _vp]),
    "goofy_b200_encode_batch_device": (_int, [_int, C.POINTER(GoofyB200Image), _u32, _vp]),
    "goofy_b200_encode_batch_sharded": (_int, [_int, C.POINTER(GoofyB200Image), _u32]),
    "goofy_b200_encode_sharded_host": (_int, [_int, _vp, _vp, _u32, _u32, _u32, _int]),
    "goofy_b200_encode_dual_sharded_host": (_int, [_vp, _vp, _vp, _u32, _u32, _u32, _int]),
    "goofy_b200_strip_partition": (None, [_u32, _int, _int, C.POINTER(_u32), C.POINTER(_u32)]),
}

_lib = None


def load() -> C.CDLL:
    """Load (building first if the sources are newer and nvcc is present).  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    import os

    from . import build as _build

    path = LIB_PATH
    override = os.environ.get("GOOFY_B200_LIB")  # experiments only: another build of the same sources
    if override:
        path = Path(override)
    elif _build.needs_build():
        _build.build_library()
    if not path.exists():
        raise ImportError(f"{path} is missing: the CUDA library is the only implementation (no CPU fallback)")
    lib = C.CDLL(str(path))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library drift: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
