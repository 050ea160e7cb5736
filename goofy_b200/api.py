"""Host-side mirror of the reference's encoder interface, on top of the C ABI.

Reference surface mirrored (GoofyTC/goofy_tc.h:10-13):
    int goofy::compressDXT1(unsigned char* result, const unsigned char* input,
                            unsigned width, unsigned height, unsigned stride);
    int goofy::compressETC1(... same ...);
Same names, same argument order and meaning (stride in BYTES, result holds width*height/2 bytes),
same return codes (0, -1 width%16, -2 height%4) -- plus the new negative codes of
include/goofy_b200.h.  Host buffers are numpy uint8 arrays (or anything exposing a writable
buffer); device buffers are torch CUDA tensors or raw integer device pointers.  torch is only
used to find a tensor's data_ptr / current stream: it is plumbing, not the encoder.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import GoofyB200Image

DXT1 = 0
ETC1 = 1
BOTH = 2   # batch entry points: DXT1 blocks to dst, ETC1s blocks to dst2, one read of every pixel
# bit-exact with goofyRef:: (Src/goofy_tc_reference.cpp) instead of with the SSE2 path
DXT1_FLOATREF = 16
ETC1_FLOATREF = 17
CODEC_NAMES = {DXT1: "dxt1", ETC1: "etc1", BOTH: "dxt1+etc1", DXT1_FLOATREF: "dxt1_floatref", ETC1_FLOATREF: "etc1_floatref"}


class GoofyError(RuntimeError):
    def __init__(self, code: int):
        self.code = code
        super().__init__(f"goofy_b200 error {code}: {error_string(code)}")


def error_string(code: int) -> str:
    return _lib.load().goofy_b200_error_string(int(code)).decode()


def device_count() -> int:
    return int(_lib.load().goofy_b200_device_count())


def kernel_launches() -> int:
    return int(_lib.load().goofy_b200_kernel_launches())


def last_launch_kernel() -> str:
    """The encode kernel this thread launched last through the library (what a call actually ran)."""
    return _lib.load().goofy_b200_last_launch_kernel().decode()


def host_scratch_sets() -> int:
    """Scratch sets of the host paths created so far (they are pooled and leased per thread)."""
    return int(_lib.load().goofy_b200_host_scratch_sets())


LOAD_AUTO, LOAD_DIRECT, LOAD_TMA, LOAD_ONESHOT, LOAD_ASYNC = 0, 1, 2, 3, 4


def set_load_path(path: int) -> int:
    """Pick the image load layer for the uniform device entry points; returns the previous setting."""
    return int(_lib.load().goofy_b200_set_load_path(path))


def get_load_path() -> int:
    return int(_lib.load().goofy_b200_get_load_path())


HOST_RGB_OFF, HOST_RGB_AUTO, HOST_RGB_ALWAYS, HOST_RGB_PAGEABLE = 0, 1, 2, 3


def set_host_rgb_staging(mode: int) -> int:
    """Alpha-stripped staging of the host path (include/goofy_b200.h); returns the previous mode."""
    return int(_lib.load().goofy_b200_set_host_rgb_staging(mode))


def get_host_rgb_staging() -> int:
    return int(_lib.load().goofy_b200_get_host_rgb_staging())


def host_threads() -> int:
    """Host threads that work on one staging job of the host path, the caller included."""
    return int(_lib.load().goofy_b200_host_threads())


def host_neighbours() -> int:
    """GPUs of this box that run somebody else's compute process (-1: cannot tell); AUTO packs pinned input only at 0."""
    return int(_lib.load().goofy_b200_host_neighbours())


def host_link_stats() -> dict:
    """Bytes the host path sent host -> device so far, and raw / alpha-stripped strips of large pinned images."""
    b, r, p, pc, nc = (C.c_uint64() for _ in range(5))
    _lib.load().goofy_b200_host_link_stats(C.byref(b), C.byref(r), C.byref(p), C.byref(pc), C.byref(nc))
    return {"bytes_uploaded": b.value, "raw_strips": r.value, "packed_strips": p.value, "packing_calls": pc.value, "plain_calls": nc.value}


def output_bytes(width: int, height: int) -> int:
    return width * height // 2


# ------------------------------------------------------------------ pointer plumbing
def _host_ptr(a, writable: bool) -> int:
    if isinstance(a, np.ndarray):
        if a.dtype != np.uint8 or not a.flags["C_CONTIGUOUS"]:
            raise TypeError("host buffers must be C-contiguous uint8 arrays")
        if writable and not a.flags["WRITEABLE"]:
            raise TypeError("result buffer is read-only")
        return a.ctypes.data
    if a is None:
        return 0
    if hasattr(a, "data_ptr"):  # torch CPU tensor (e.g. pinned)
        if a.is_cuda:
            raise TypeError("device tensor passed to the host API")
        return int(a.data_ptr())
    return int(a)


def _dev_ptr(t) -> int:
    if t is None:
        return 0
    if hasattr(t, "data_ptr"):
        if not t.is_cuda:
            raise TypeError("host tensor passed to the device API")
        return int(t.data_ptr())
    return int(t)


def _stream_ptr(stream) -> int:
    if stream is None:
        try:
            import torch

            if torch.cuda.is_available():
                return int(torch.cuda.current_stream().cuda_stream)
        except ImportError:
            pass
        return 0
    if hasattr(stream, "cuda_stream"):
        return int(stream.cuda_stream)
    return int(stream)


# ------------------------------------------------------------------ drop-in host API
def compressDXT1(result, input, width: int, height: int, stride: int) -> int:
    """goofy::compressDXT1 (GoofyTC/goofy_tc.h:1497).  Returns the status code; never raises on bad shapes."""
    return int(_lib.load().goofy_b200_compress_dxt1(_host_ptr(result, True), _host_ptr(input, False),
                                                    width, height, stride))


def compressETC1(result, input, width: int, height: int, stride: int) -> int:
    """goofy::compressETC1 (GoofyTC/goofy_tc.h:1528)."""
    return int(_lib.load().goofy_b200_compress_etc1(_host_ptr(result, True), _host_ptr(input, False),
                                                    width, height, stride))


class goofyRef:
    """Mirror of namespace goofyRef (Src/goofy_tc_reference.h:5-9): the float-reference flavour."""

    @staticmethod
    def compressDXT1(result, input, width: int, height: int, stride: int) -> int:
        return int(_lib.load().goofy_b200_compress_dxt1_floatref(_host_ptr(result, True), _host_ptr(input, False),
                                                                 width, height, stride))

    @staticmethod
    def compressETC1(result, input, width: int, height: int, stride: int) -> int:
        return int(_lib.load().goofy_b200_compress_etc1_floatref(_host_ptr(result, True), _host_ptr(input, False),
                                                                 width, height, stride))


def encode_host(codec: int, result, input, width: int, height: int, stride: int) -> int:
    return int(_lib.load().goofy_b200_encode_host(codec, _host_ptr(result, True), _host_ptr(input, False),
                                                  width, height, stride))


def encode_dual_host(result_dxt1, result_etc1, input, width: int, height: int, stride: int) -> int:
    """DXT1 and ETC1s of one host image from a single upload (4 B/px over PCIe instead of 8)."""
    return int(_lib.load().goofy_b200_encode_dual_host(_host_ptr(result_dxt1, True), _host_ptr(result_etc1, True),
                                                       _host_ptr(input, False), width, height, stride))


def encode_rgb24_host(codec: int, result, input, width: int, height: int, stride: int, result2=None) -> int:
    """A host image of packed RGB8 pixels (3 bytes per pixel, rows `stride` >= width*3 bytes apart): 3 instead of 4 input
    bytes per pixel cross PCIe.  codec may be BOTH (result2 = the ETC1s blocks)."""
    return int(_lib.load().goofy_b200_encode_rgb24_host(codec, _host_ptr(result, True), _host_ptr(result2, True) if result2 is not None else 0,
                                                        _host_ptr(input, False), width, height, stride))


def encode_host_batch(codec: int, images) -> int:
    """images: iterable of (input, result, width, height, stride[, result2]) with HOST buffers (numpy uint8 arrays or pinned
    torch tensors); codec BOTH needs result2 (the ETC1s blocks).  One pipeline for all of them: copies and kernels of
    neighbouring images overlap."""
    items = list(images)
    arr = (GoofyB200Image * max(len(items), 1))()
    for i, it in enumerate(items):
        src, dst, w, h, stride = it[:5]
        dst2 = it[5] if len(it) > 5 else None
        arr[i] = GoofyB200Image(_host_ptr(src, False), _host_ptr(dst, True), w, h, stride, -1, _host_ptr(dst2, True))
    return int(_lib.load().goofy_b200_encode_host_batch(codec, arr, len(items)))


def encode_sharded_host(codec: int, result, input, width: int, height: int, stride: int, n_gpus: int = 0) -> int:
    """One host image, horizontal strips of whole block rows, strip g on GPU g (no collectives)."""
    return int(_lib.load().goofy_b200_encode_sharded_host(codec, _host_ptr(result, True), _host_ptr(input, False),
                                                          width, height, stride, n_gpus))


def encode_dual_sharded_host(result_dxt1, result_etc1, input, width: int, height: int, stride: int, n_gpus: int = 0) -> int:
    """The same partition, both codecs from one upload of every strip."""
    return int(_lib.load().goofy_b200_encode_dual_sharded_host(_host_ptr(result_dxt1, True), _host_ptr(result_etc1, True),
                                                               _host_ptr(input, False), width, height, stride, n_gpus))


# ------------------------------------------------------------------ device-resident API
def encode_device(codec: int, d_result, d_input, width: int, height: int, stride: int, stream=None) -> int:
    """Asynchronous on `stream` (default: torch's current stream)."""
    return int(_lib.load().goofy_b200_encode_device(codec, _dev_ptr(d_result), _dev_ptr(d_input), width, height,
                                                    stride, _stream_ptr(stream)))


def encode_rgb24_device(codec: int, d_result, d_input, width: int, height: int, stride: int, d_result2=None,
                        input_image_pitch: int = 0, result_image_pitch: int = 0, n_images: int = 1, stream=None) -> int:
    """Packed RGB8 input (3 bytes per pixel, rows `stride` >= width*3 bytes apart, 4-byte aligned): the same bytes the
    RGBA entry points produce for the same pixels.  codec may be BOTH (d_result2 = the ETC1s blocks)."""
    return int(_lib.load().goofy_b200_encode_rgb24_device(
        codec, _dev_ptr(d_result), _dev_ptr(d_result2), _dev_ptr(d_input), width, height, stride, input_image_pitch,
        result_image_pitch, n_images, _stream_ptr(stream)))


def encode_relaxed_device(codec: int, d_result, d_input, width: int, height: int, stride: int, stream=None) -> int:
    """Any width / height (edge blocks replicate the last column / row); result holds ceil(w/4)*ceil(h/4) blocks."""
    return int(_lib.load().goofy_b200_encode_relaxed_device(codec, _dev_ptr(d_result), _dev_ptr(d_input), width, height,
                                                            stride, _stream_ptr(stream)))


def encode_batch_uniform_device(codec: int, d_result, d_input, width: int, height: int, stride: int,
                                input_image_pitch: int, result_image_pitch: int, n_images: int, stream=None) -> int:
    return int(_lib.load().goofy_b200_encode_batch_uniform_device(
        codec, _dev_ptr(d_result), _dev_ptr(d_input), width, height, stride, input_image_pitch,
        result_image_pitch, n_images, _stream_ptr(stream)))


def encode_dual_device(d_result_dxt1, d_result_etc1, d_input, width: int, height: int, stride: int,
                       input_image_pitch: int = 0, result_image_pitch: int = 0, n_images: int = 1, stream=None) -> int:
    """DXT1 and ETC1s from one read of the input."""
    return int(_lib.load().goofy_b200_encode_dual_device(
        _dev_ptr(d_result_dxt1), _dev_ptr(d_result_etc1), _dev_ptr(d_input), width, height, stride,
        input_image_pitch, result_image_pitch, n_images, _stream_ptr(stream)))


def decode_device(codec: int, d_rgba, d_blocks, width: int, height: int, stride: int, stream=None) -> int:
    """BC1 / ETC1 blocks -> RGBA8 on the device (DecoderBC::decodeBlockDXT1/ETC1, Src/decoder.cpp:933-971)."""
    return int(_lib.load().goofy_b200_decode_device(codec, _dev_ptr(d_rgba), _dev_ptr(d_blocks), width, height, stride,
                                                    _stream_ptr(stream)))


def block_sse_device(codec: int, d_blocks, d_rgba, width: int, height: int, stride: int, d_sse_rgb, stream=None) -> int:
    """Adds per-channel sums of squared (decoded - source) to the 3 x uint64 device array d_sse_rgb."""
    return int(_lib.load().goofy_b200_block_sse_device(codec, _dev_ptr(d_blocks), _dev_ptr(d_rgba), width, height, stride,
                                                       _dev_ptr(d_sse_rgb), _stream_ptr(stream)))


def psnr_rgb768(sse_rgb, pixels: int) -> float:
    """The reference harness's RGB-PSNR (Src/main.cpp:444,466): peak 768 over the SUM of channel MSEs."""
    import math
    mse = float(sum(int(v) for v in sse_rgb)) / pixels
    return float("inf") if mse < 1e-7 else 10.0 * math.log10(768.0 * 768.0 / mse)


def make_descriptors(images: Iterable[Sequence]) -> C.Array:
    """images: iterable of (d_src, d_dst, width, height, stride[, device[, d_dst2]]) -> GoofyB200Image[n]."""
    items = list(images)
    arr = (GoofyB200Image * len(items))()
    for i, it in enumerate(items):
        src, dst, w, h, stride = it[:5]
        dev = it[5] if len(it) > 5 else -1
        dst2 = it[6] if len(it) > 6 else None
        arr[i] = GoofyB200Image(_dev_ptr(src), _dev_ptr(dst), w, h, stride, dev, _dev_ptr(dst2))
    return arr


def encode_batch_device(codec: int, images, stream=None) -> int:
    """Ragged batch on the current device, one launch."""
    descs = images if isinstance(images, C.Array) else make_descriptors(images)
    return int(_lib.load().goofy_b200_encode_batch_device(codec, descs, len(descs), _stream_ptr(stream)))


def encode_batch_sharded(codec: int, images) -> int:
    """Ragged batch whose images live on several GPUs of this process; returns after all devices finish."""
    descs = images if isinstance(images, C.Array) else make_descriptors(images)
    return int(_lib.load().goofy_b200_encode_batch_sharded(codec, descs, len(descs)))


def strip_partition(height: int, n_shards: int, shard: int) -> tuple[int, int]:
    """(first_block_row, block_row_count) of `shard` -- the scheduler's own partition function."""
    first, count = C.c_uint32(), C.c_uint32()
    _lib.load().goofy_b200_strip_partition(height, n_shards, shard, C.byref(first), C.byref(count))
    return first.value, count.value


def check(code: int) -> None:
    if code != 0:
        raise GoofyError(code)
