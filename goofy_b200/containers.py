"""On-disk formats either side of the encoders -- host-side glue, no CUDA.

* ``write_dds`` / ``write_ktx``: containers the reference harness emits for its results
  (saveDds Src/main.cpp:154-183, FourCC 'DXT1' :72; saveKtx :185-220, GL_ETC1_RGB8_OES 0x8D64 :65) so the
  blocks are loadable by standard tools.  Field values follow the DDS / KTX 1.1 specifications and match
  what the reference writes (one mip level, one face, linear size = payload bytes).
* ``read_dds`` / ``read_ktx``: the inverse, used by the tests.
* ``load_png_rgba``: PNG ingest with the reference loader's contract (loadPngAsRgba8, Src/main.cpp:258-343):
  RGBA8, alpha forced to 255, width % 16 == 0 and height % 4 == 0 or the image is rejected, 64-byte aligned.
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

DDS_MAGIC = 0x20534444          # "DDS "
FOURCC_DXT1 = 0x31545844        # "DXT1"
KTX_IDENTIFIER = bytes([0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A])
KTX_ENDIANNESS = 0x04030201
GL_ETC1_RGB8_OES = 0x8D64
GL_COMPRESSED_RGB_S3TC_DXT1_EXT = 0x83F0
GL_RGB, GL_RGBA = 0x1907, 0x1908

# DDSD_CAPS | DDSD_HEIGHT | DDSD_WIDTH | DDSD_PIXELFORMAT | DDSD_MIPMAPCOUNT | DDSD_LINEARSIZE
DDS_FLAGS = 0x1 | 0x2 | 0x4 | 0x1000 | 0x20000 | 0x80000
DDPF_FOURCC = 0x4


def _payload(blocks, width: int, height: int) -> bytes:
    data = np.ascontiguousarray(blocks, dtype=np.uint8).tobytes()
    if len(data) != width * height // 2:
        raise ValueError(f"expected {width * height // 2} bytes of 8-byte blocks, got {len(data)}")
    return data


def dds_bytes(blocks, width: int, height: int, fourcc: int = FOURCC_DXT1) -> bytes:
    data = _payload(blocks, width, height)
    pixel_format = struct.pack("<8I", 32, DDPF_FOURCC, fourcc, 0, 0, 0, 0, 0)
    header = struct.pack("<8I", DDS_MAGIC, 124, DDS_FLAGS, height, width, len(data), 1, 1)
    header += b"\0" * 44 + pixel_format + struct.pack("<5I", 0, 0, 0, 0, 0)
    assert len(header) == 128
    return header + data


def write_dds(path, blocks, width: int, height: int, fourcc: int = FOURCC_DXT1) -> None:
    Path(path).write_bytes(dds_bytes(blocks, width, height, fourcc))


def read_dds(path):
    raw = Path(path).read_bytes()
    magic, size, flags, height, width, linear, depth, mips = struct.unpack_from("<8I", raw, 0)
    pf_size, pf_flags, fourcc = struct.unpack_from("<3I", raw, 76)
    if magic != DDS_MAGIC or size != 124 or pf_size != 32 or not (pf_flags & DDPF_FOURCC):
        raise ValueError("not a FourCC DDS file")
    return {"width": width, "height": height, "fourcc": fourcc, "mips": mips, "linear_size": linear,
            "blocks": np.frombuffer(raw, dtype=np.uint8, offset=128, count=linear).copy()}


def ktx_bytes(blocks, width: int, height: int, gl_internal_format: int = GL_ETC1_RGB8_OES) -> bytes:
    data = _payload(blocks, width, height)
    # the reference's rule (saveKtx, Src/main.cpp:199): GL_RGB for ETC1 only, GL_RGBA for every other format
    base = GL_RGB if gl_internal_format == GL_ETC1_RGB8_OES else GL_RGBA
    header = KTX_IDENTIFIER + struct.pack("<13I", KTX_ENDIANNESS, 0, 1, 0, gl_internal_format, base, width, height,
                                          0, 0, 1, 1, 0)
    assert len(header) == 64
    return header + struct.pack("<I", len(data)) + data


def write_ktx(path, blocks, width: int, height: int, gl_internal_format: int = GL_ETC1_RGB8_OES) -> None:
    Path(path).write_bytes(ktx_bytes(blocks, width, height, gl_internal_format))


def read_ktx(path):
    raw = Path(path).read_bytes()
    if raw[:12] != KTX_IDENTIFIER:
        raise ValueError("not a KTX 1.1 file")
    (endian, gl_type, type_size, gl_format, internal, base, width, height, depth, elements, faces, mips,
     kv_bytes) = struct.unpack_from("<13I", raw, 12)
    if endian != KTX_ENDIANNESS:
        raise ValueError("big-endian KTX not supported")
    (image_size,) = struct.unpack_from("<I", raw, 64 + kv_bytes)
    return {"width": width, "height": height, "gl_internal_format": internal, "gl_base_internal_format": base,
            "mips": mips, "faces": faces,
            "blocks": np.frombuffer(raw, dtype=np.uint8, offset=68 + kv_bytes, count=image_size).copy()}


def load_png_rgba(path, align: int = 64) -> np.ndarray:
    """RGBA8 [h, w, 4], alpha 255, contiguous and `align`-byte aligned; ValueError if the encoders cannot take it."""
    from PIL import Image

    with Image.open(path) as im:
        rgba = np.array(im.convert("RGBA"), dtype=np.uint8)
    h, w = rgba.shape[:2]
    if w % 16 != 0:
        raise ValueError(f"{path}: width {w} is not a multiple of 16")
    if h % 4 != 0:
        raise ValueError(f"{path}: height {h} is not a multiple of 4")
    rgba[..., 3] = 255
    raw = np.empty(rgba.size + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    out = raw[off:off + rgba.size].reshape(h, w, 4)
    out[...] = rgba
    return out
