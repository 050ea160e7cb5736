"""goofy_b200 -- B200 (sm_100a) implementation of Goofy's DXT1/BC1 and ETC1s block encoders.

Public surface (mirrors GoofyTC/goofy_tc.h:10-13 plus the batched device-resident variants):
    compressDXT1, compressETC1                       drop-in host API, same signature and codes
    encode_host_batch                                many host images through one pipeline
    encode_device, encode_batch_uniform_device,
    encode_dual_device, encode_batch_device          device-resident, asynchronous
    encode_sharded_host, encode_batch_sharded        multi-GPU, no collectives
The implementation is libgoofy_b200.so (CUDA only; there is no CPU fallback).
"""
from .api import (encode_rgb24_host, encode_rgb24_device, set_host_rgb_staging, get_host_rgb_staging, host_threads, host_neighbours, host_link_stats,
                  HOST_RGB_OFF, HOST_RGB_AUTO, HOST_RGB_ALWAYS, HOST_RGB_PAGEABLE, encode_relaxed_device, DXT1_FLOATREF, ETC1_FLOATREF, goofyRef, decode_device, block_sse_device, psnr_rgb768, DXT1, ETC1, BOTH, CODEC_NAMES, GoofyError, check, compressDXT1, compressETC1, device_count,
                  encode_batch_device, encode_batch_sharded, encode_batch_uniform_device, encode_device,
                  encode_dual_device, encode_dual_host, encode_host, encode_host_batch, encode_sharded_host, encode_dual_sharded_host, error_string, kernel_launches, host_scratch_sets, last_launch_kernel,
                  make_descriptors, output_bytes, strip_partition, set_load_path, get_load_path, LOAD_AUTO,
                  LOAD_DIRECT, LOAD_TMA, LOAD_ONESHOT, LOAD_ASYNC)

__all__ = [
    "encode_rgb24_host", "encode_rgb24_device", "set_host_rgb_staging", "get_host_rgb_staging", "host_threads", "host_neighbours", "host_link_stats",
    "HOST_RGB_OFF", "HOST_RGB_AUTO", "HOST_RGB_ALWAYS", "HOST_RGB_PAGEABLE",
    "encode_relaxed_device",
    "DXT1_FLOATREF", "ETC1_FLOATREF", "goofyRef",
    "decode_device", "block_sse_device", "psnr_rgb768",
    "DXT1", "ETC1", "BOTH", "CODEC_NAMES", "GoofyError", "check", "compressDXT1", "compressETC1", "device_count",
    "encode_batch_device", "encode_batch_sharded", "encode_batch_uniform_device", "encode_device",
    "encode_dual_device", "encode_dual_host", "encode_host", "encode_host_batch", "encode_sharded_host", "encode_dual_sharded_host", "error_string", "kernel_launches", "host_scratch_sets", "last_launch_kernel",
    "make_descriptors", "output_bytes", "strip_partition", "set_load_path", "get_load_path", "LOAD_AUTO",
    "LOAD_DIRECT", "LOAD_TMA", "LOAD_ONESHOT", "LOAD_ASYNC",
]
