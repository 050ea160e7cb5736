// capi.cu -- the extern "C" layer of libgoofy_b200.so (declared in include/goofy_b200.h).
//
// Thin by design: the entry points only validate and dispatch.  The host side is split into
//   host_common.cuh    error codes, per-device set-up, the reference's argument checks
//   host_launch.cuh    load-layer policy and kernel launches (device-resident entry points)
//   host_resources.cuh per-thread scratch of the host paths, leased from a process-wide pool
//   host_pipeline.cuh  the drop-in host-pointer path (strip pipeline, pinned staging)
//   host_batch.cuh     ragged batches and the multi-GPU shard scheduler
// and everything is compiled as ONE translation unit (the kernels' __device__ tables live in it).
// All arithmetic lives in block_codec.cuh.  There is no CPU encoder anywhere in this library:
// without a CUDA device every call returns an error code.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/goofy_b200.h"
#include "encode_kernels.cuh"
#include "tma_kernels.cuh"
#include "decode_kernels.cuh"

#include "host_common.cuh"
#include "host_launch.cuh"
#include "host_pipeline.cuh"
#include "host_batch.cuh"


// ======================================================================== extern "C"
extern "C" {

int goofy_b200_abi_version(void) { return GOOFY_B200_ABI_VERSION; }

int goofy_b200_device_count(void) { return device_count(); }

uint64_t goofy_b200_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

const char* goofy_b200_last_launch_kernel(void) { return t_lastKernel; }

uint64_t goofy_b200_host_scratch_sets(void) { return ResourcePool::get().created(); }

int goofy_b200_set_load_path(int path)
{
    if (path < GOOFY_B200_LOAD_AUTO || path > GOOFY_B200_LOAD_ASYNC) return GOOFY_B200_E_ARGS;
    return g_loadPath.exchange(path, std::memory_order_relaxed);
}

int goofy_b200_get_load_path(void) { return g_loadPath.load(std::memory_order_relaxed); }

int goofy_b200_set_host_rgb_staging(int mode)
{
    if (mode < GOOFY_B200_HOST_RGB_OFF || mode > GOOFY_B200_HOST_RGB_PAGEABLE) return GOOFY_B200_E_ARGS;
    const int before = host_rgb_mode();
    g_hostRgb.store(mode, std::memory_order_relaxed);
    return before;
}

int goofy_b200_get_host_rgb_staging(void) { return host_rgb_mode(); }

int goofy_b200_host_threads(void) { return (int)CopyPool::get().threads(); }

int goofy_b200_host_neighbours(void) { return Neighbours::count(); }

void goofy_b200_host_link_stats(uint64_t* bytes_uploaded, uint64_t* raw_strips, uint64_t* packed_strips, uint64_t* packing_calls,
                                uint64_t* plain_calls)
{
    if (packing_calls) *packing_calls = g_packingCalls.load(std::memory_order_relaxed);
    if (plain_calls) *plain_calls = g_plainCalls.load(std::memory_order_relaxed);
    if (bytes_uploaded) *bytes_uploaded = g_hostUploaded.load(std::memory_order_relaxed);
    if (raw_strips) *raw_strips = g_rawStrips.load(std::memory_order_relaxed);
    if (packed_strips) *packed_strips = g_packedStrips.load(std::memory_order_relaxed);
}

const char* goofy_b200_error_string(int code)
{
    switch (code) {
        case GOOFY_B200_OK: return "ok";
        case GOOFY_B200_E_WIDTH: return "width is not a multiple of 16";
        case GOOFY_B200_E_HEIGHT: return "height is not a multiple of 4";
        case GOOFY_B200_E_NULL: return "null pointer";
        case GOOFY_B200_E_ALIGN: return "input/stride must be 16-byte aligned, output 8-byte aligned";
        case GOOFY_B200_E_STRIDE: return "stride smaller than width*4";
        case GOOFY_B200_E_CODEC: return "unknown codec";
        case GOOFY_B200_E_DEVICE: return "no usable CUDA device";
        case GOOFY_B200_E_ARGS: return "invalid argument";
        default: break;
    }
    if (code <= GOOFY_B200_E_CUDA_BASE) return cudaGetErrorString((cudaError_t)(GOOFY_B200_E_CUDA_BASE - code));
    return "unknown error";
}

int goofy_b200_compress_dxt1(unsigned char* result, const unsigned char* input, unsigned int width, unsigned int height,
                             unsigned int stride)
{
    return encode_host(GOOFY_B200_DXT1, result, input, width, height, stride);
}

int goofy_b200_compress_etc1(unsigned char* result, const unsigned char* input, unsigned int width, unsigned int height,
                             unsigned int stride)
{
    return encode_host(GOOFY_B200_ETC1, result, input, width, height, stride);
}

int goofy_b200_encode_host(int codec, void* result, const void* input, uint32_t width, uint32_t height, uint32_t stride)
{
    return encode_host(codec, result, input, width, height, stride);
}

int goofy_b200_encode_dual_host(void* result_dxt1, void* result_etc1, const void* input, uint32_t width, uint32_t height,
                                uint32_t stride)
{
    return encode_dual_host(result_dxt1, result_etc1, input, width, height, stride);
}

int goofy_b200_encode_rgb24_host(int codec, void* result, void* result2, const void* input, uint32_t width, uint32_t height,
                                 uint32_t stride)
{
    return encode_rgb24_host(codec, result, result2, input, width, height, stride);
}

int goofy_b200_encode_host_batch(int codec, const GoofyB200Image* images, uint32_t n_images)
{
    return encode_host_batch(codec, images, n_images);
}

int goofy_b200_encode_device(int codec, void* d_result, const void* d_input, uint32_t width, uint32_t height,
                             uint32_t stride, void* stream)
{
    return encode_any(codec, d_result, d_input, width, height, stride, 0, 0, 1, (cudaStream_t)stream);
}

int goofy_b200_compress_dxt1_floatref(unsigned char* result, const unsigned char* input, unsigned int width,
                                      unsigned int height, unsigned int stride)
{
    return encode_host(GOOFY_B200_DXT1_FLOATREF, result, input, width, height, stride);
}

int goofy_b200_compress_etc1_floatref(unsigned char* result, const unsigned char* input, unsigned int width,
                                      unsigned int height, unsigned int stride)
{
    return encode_host(GOOFY_B200_ETC1_FLOATREF, result, input, width, height, stride);
}

int goofy_b200_encode_batch_uniform_device(int codec, void* d_result, const void* d_input, uint32_t width, uint32_t height,
                                           uint32_t stride, uint64_t input_image_pitch, uint64_t result_image_pitch,
                                           uint32_t n_images, void* stream)
{
    return encode_any(codec, d_result, d_input, width, height, stride, input_image_pitch, result_image_pitch, n_images,
                      (cudaStream_t)stream);
}

int goofy_b200_encode_dual_device(void* d_result_dxt1, void* d_result_etc1, const void* d_input, uint32_t width,
                                  uint32_t height, uint32_t stride, uint64_t input_image_pitch, uint64_t result_image_pitch,
                                  uint32_t n_images, void* stream)
{
    return encode_uniform(gb::kDual, d_result_dxt1, d_result_etc1, d_input, width, height, stride, input_image_pitch,
                          result_image_pitch, n_images, (cudaStream_t)stream);
}

int goofy_b200_encode_rgb24_device(int codec, void* d_result, void* d_result2, const void* d_input, uint32_t width, uint32_t height,
                                   uint32_t stride, uint64_t input_image_pitch, uint64_t result_image_pitch, uint32_t n_images,
                                   void* stream)
{
    return encode_rgb24(codec, d_result, d_result2, d_input, width, height, stride, input_image_pitch, result_image_pitch, n_images,
                        (cudaStream_t)stream);
}

int goofy_b200_encode_relaxed_device(int codec, void* d_result, const void* d_input, uint32_t width, uint32_t height,
                                     uint32_t stride, void* stream)
{
    if (!is_codec(codec)) return GOOFY_B200_E_CODEC;
    if (width == 0u || height == 0u) return GOOFY_B200_OK;
    if ((uint64_t)stride < (uint64_t)width * 4u) return GOOFY_B200_E_STRIDE;
    if (!d_result || !d_input) return GOOFY_B200_E_NULL;
    if (((uintptr_t)d_input & 3u) != 0u || (stride & 3u) != 0u || ((uintptr_t)d_result & 7u) != 0u) return GOOFY_B200_E_ALIGN;
    const uint32_t bw = (width + 3u) / 4u, bh = (height + 3u) / 4u;
    int rc = ensure_device_ready();
    if (rc != GOOFY_B200_OK) return rc;
    // DXT1 flavours: one block row per CTA; ETC1s flavours: four (their CTAs stage the control table first)
    const bool etc = codec == GOOFY_B200_ETC1 || codec == GOOFY_B200_ETC1_FLOATREF;
    uint32_t gy = etc ? (bh + 3u) / 4u : bh;
    if (gy > 65535u) gy = 65535u;   // taller images: every CTA simply walks further
    const dim3 grid((bw + 255u) / 256u, gy, 1);
    cudaStream_t s = (cudaStream_t)stream;
    const uint8_t* src = (const uint8_t*)d_input;
    uint8_t* dst = (uint8_t*)d_result;
    t_lastKernel = "encode_relaxed_kernel";
    const dim3 block(256, 1, 1);
    switch (codec) {
        case GOOFY_B200_DXT1: return launch_pdl(gb::encode_relaxed_kernel<gb::kDxt1, 0>, grid, block, s, src, dst, width, height, stride);
        case GOOFY_B200_ETC1: return launch_pdl(gb::encode_relaxed_kernel<gb::kEtc1, 0>, grid, block, s, src, dst, width, height, stride);
        case GOOFY_B200_DXT1_FLOATREF: return launch_pdl(gb::encode_relaxed_kernel<gb::kDxt1, 1>, grid, block, s, src, dst, width, height, stride);
        default: return launch_pdl(gb::encode_relaxed_kernel<gb::kEtc1, 1>, grid, block, s, src, dst, width, height, stride);
    }
}

int goofy_b200_decode_device(int codec, void* d_rgba, const void* d_blocks, uint32_t width, uint32_t height, uint32_t stride,
                             void* stream)
{
    if (codec != GOOFY_B200_DXT1 && codec != GOOFY_B200_ETC1) return GOOFY_B200_E_CODEC;
    if (width % 4u != 0u) return GOOFY_B200_E_WIDTH;
    if (height % 4u != 0u) return GOOFY_B200_E_HEIGHT;
    if (width == 0u || height == 0u) return GOOFY_B200_OK;
    if ((uint64_t)stride < (uint64_t)width * 4u) return GOOFY_B200_E_STRIDE;
    if (!d_rgba || !d_blocks) return GOOFY_B200_E_NULL;
    if (((uintptr_t)d_rgba & 15u) != 0u || (stride & 15u) != 0u || ((uintptr_t)d_blocks & 7u) != 0u) return GOOFY_B200_E_ALIGN;
    if (height / 4u > 65535u) return GOOFY_B200_E_ARGS;
    int rc = ensure_device_ready();
    if (rc != GOOFY_B200_OK) return rc;
    gb::DecodeParams P;
    P.blocks = (const uint8_t*)d_blocks;
    P.rgba = (uint8_t*)d_rgba;
    P.source = nullptr;
    P.sse = nullptr;
    P.bw = width / 4u;
    P.bh = height / 4u;
    P.stride = stride;
    const dim3 grid((P.bw + 255u) / 256u, P.bh, 1);
    return codec == GOOFY_B200_DXT1 ? launch_pdl(gb::decode_kernel<0>, grid, dim3(256, 1, 1), (cudaStream_t)stream, P)
                                    : launch_pdl(gb::decode_kernel<1>, grid, dim3(256, 1, 1), (cudaStream_t)stream, P);
}

int goofy_b200_block_sse_device(int codec, const void* d_blocks, const void* d_rgba, uint32_t width, uint32_t height,
                                uint32_t stride, uint64_t* d_sse_rgb, void* stream)
{
    if (codec != GOOFY_B200_DXT1 && codec != GOOFY_B200_ETC1) return GOOFY_B200_E_CODEC;
    if (width % 4u != 0u) return GOOFY_B200_E_WIDTH;
    if (height % 4u != 0u) return GOOFY_B200_E_HEIGHT;
    if (width == 0u || height == 0u) return GOOFY_B200_OK;
    if ((uint64_t)stride < (uint64_t)width * 4u) return GOOFY_B200_E_STRIDE;
    if (!d_rgba || !d_blocks || !d_sse_rgb) return GOOFY_B200_E_NULL;
    if (((uintptr_t)d_rgba & 15u) != 0u || (stride & 15u) != 0u || ((uintptr_t)d_blocks & 7u) != 0u ||
        ((uintptr_t)d_sse_rgb & 7u) != 0u)
        return GOOFY_B200_E_ALIGN;
    if (height / 4u > 65535u) return GOOFY_B200_E_ARGS;
    int rc = ensure_device_ready();
    if (rc != GOOFY_B200_OK) return rc;
    gb::DecodeParams P;
    P.blocks = (const uint8_t*)d_blocks;
    P.rgba = nullptr;
    P.source = (const uint8_t*)d_rgba;
    P.sse = (unsigned long long*)d_sse_rgb;
    P.bw = width / 4u;
    P.bh = height / 4u;
    P.stride = stride;
    // CTAs walk down the image: about eight resident waves of them (measured 6374 / 6509 / 6465 / 6643 / 6390 GB/s at
    // 1 / 2 / 4 / 8 / 16 waves on one 8192^2 texture: CTAs that retire at staggered times beat fully persistent ones,
    // one block row per CTA pays too many atomics), and never more than kSseMaxBlocksPerThread block rows per thread
    // (32-bit partial sums)
    const uint32_t gx = (P.bw + 255u) / 256u;
    int dev = -1;
    cudaGetDevice(&dev);
    const int sms = sm_count(dev);
    static const uint32_t waves = []() { const char* e = getenv("GOOFY_B200_SSE_WAVES"); const int v = e ? atoi(e) : 8; return v > 0 ? (uint32_t)v : 8u; }();
    uint32_t gy = (uint32_t)(sms > 0 ? sms : 148) * 8u * waves / gx;
    const uint32_t gyMin = (P.bh + gb::kSseMaxBlocksPerThread - 1u) / gb::kSseMaxBlocksPerThread;
    if (gy < gyMin) gy = gyMin;
    if (gy < 1u) gy = 1u;
    if (gy > P.bh) gy = P.bh;
    const dim3 grid(gx, gy, 1);
    return codec == GOOFY_B200_DXT1 ? launch_pdl(gb::block_sse_kernel<0>, grid, dim3(256, 1, 1), (cudaStream_t)stream, P)
                                    : launch_pdl(gb::block_sse_kernel<1>, grid, dim3(256, 1, 1), (cudaStream_t)stream, P);
}

int goofy_b200_encode_batch_device(int codec, const GoofyB200Image* descs, uint32_t n_images, void* stream)
{
    if (n_images && descs) {
        int dev = -1;
        if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return GOOFY_B200_E_DEVICE; }
        for (uint32_t i = 0; i < n_images; ++i)
            if (descs[i].device >= 0 && descs[i].device != dev) return GOOFY_B200_E_DEVICE;
    }
    return encode_batch_current_device(codec, descs, nullptr, n_images, (cudaStream_t)stream);
}

int goofy_b200_encode_batch_sharded(int codec, const GoofyB200Image* descs, uint32_t n_images)
{
    if (codec != GOOFY_B200_BOTH && !is_codec(codec)) return GOOFY_B200_E_CODEC;
    if (n_images == 0u) return GOOFY_B200_OK;
    if (!descs) return GOOFY_B200_E_NULL;
    const int nDev = device_count();
    if (nDev <= 0) return GOOFY_B200_E_DEVICE;
    std::vector<std::vector<uint32_t>> perDevice((size_t)nDev);
    for (uint32_t i = 0; i < n_images; ++i) {
        if (descs[i].device < 0 || descs[i].device >= nDev) return GOOFY_B200_E_DEVICE;
        perDevice[(size_t)descs[i].device].push_back(i);
    }
    std::lock_guard<std::mutex> lock(g_schedMutex);
    std::vector<int> used;
    for (int d = 0; d < nDev; ++d) {
        if (perDevice[(size_t)d].empty()) continue;
        const std::vector<uint32_t>* order = &perDevice[(size_t)d];
        worker_for(d)->submit([codec, descs, order]() {
            int rc = encode_batch_current_device(codec, descs, order->data(), (uint32_t)order->size(), nullptr);
            if (rc != GOOFY_B200_OK) return rc;
            return cuda_rc(cudaStreamSynchronize(nullptr));
        });
        used.push_back(d);
    }
    int rc = GOOFY_B200_OK;
    for (int d : used) {
        const int r = g_workers[(size_t)d]->wait();
        if (r != GOOFY_B200_OK && rc == GOOFY_B200_OK) rc = r;
    }
    return rc;
}

void goofy_b200_strip_partition(uint32_t height, int n_shards, int shard, uint32_t* first_block_row,
                                uint32_t* block_row_count)
{
    const uint32_t rows = height / 4u;
    uint32_t first = 0, count = 0;
    if (n_shards > 0 && shard >= 0 && shard < n_shards) {
        first = (uint32_t)((uint64_t)rows * (uint32_t)shard / (uint32_t)n_shards);
        const uint32_t next = (uint32_t)((uint64_t)rows * ((uint32_t)shard + 1u) / (uint32_t)n_shards);
        count = next - first;
    }
    if (first_block_row) *first_block_row = first;
    if (block_row_count) *block_row_count = count;
}

// One host image in strips of whole block rows, strip g on device g through the host path (result2 != null: both codecs).
static int sharded_host(int codec, void* result, void* result2, const void* input, uint32_t width, uint32_t height, uint32_t stride,
                        int n_gpus)
{
    int rc = check_shape(width, height, stride);
    if (rc != GOOFY_B200_OK) return rc;
    if (width == 0u || height == 0u) return GOOFY_B200_OK;
    if (!result || !input) return GOOFY_B200_E_NULL;
    const int nDev = device_count();
    if (nDev <= 0) return GOOFY_B200_E_DEVICE;
    if (n_gpus <= 0) n_gpus = nDev;
    if (n_gpus > nDev) return GOOFY_B200_E_DEVICE;

    std::lock_guard<std::mutex> lock(g_schedMutex);
    std::vector<int> used;
    for (int g = 0; g < n_gpus; ++g) {
        uint32_t first = 0, count = 0;
        goofy_b200_strip_partition(height, n_gpus, g, &first, &count);
        if (count == 0u) continue;
        const size_t outOffset = (size_t)first * (width / 4u) * 8u;
        uint8_t* out = (uint8_t*)result + outOffset;
        uint8_t* out2 = result2 ? (uint8_t*)result2 + outOffset : nullptr;
        const uint8_t* in = (const uint8_t*)input + (size_t)first * 4u * stride;
        worker_for(g)->submit([=]() {
            return out2 ? encode_dual_host(out, out2, in, width, count * 4u, stride) : encode_host(codec, out, in, width, count * 4u, stride);
        });
        used.push_back(g);
    }
    for (int g : used) {
        const int r = g_workers[(size_t)g]->wait();
        if (r != GOOFY_B200_OK && rc == GOOFY_B200_OK) rc = r;
    }
    return rc;
}

int goofy_b200_encode_sharded_host(int codec, void* result, const void* input, uint32_t width, uint32_t height,
                                   uint32_t stride, int n_gpus)
{
    if (codec != GOOFY_B200_DXT1 && codec != GOOFY_B200_ETC1) return GOOFY_B200_E_CODEC;
    return sharded_host(codec, result, nullptr, input, width, height, stride, n_gpus);
}

int goofy_b200_encode_dual_sharded_host(void* result_dxt1, void* result_etc1, const void* input, uint32_t width, uint32_t height,
                                        uint32_t stride, int n_gpus)
{
    if (width != 0u && height != 0u && width % 16u == 0u && height % 4u == 0u && !result_etc1) return GOOFY_B200_E_NULL;
    return sharded_host(GOOFY_B200_DXT1, result_dxt1, result_etc1, input, width, height, stride, n_gpus);
}

}  // extern "C"
