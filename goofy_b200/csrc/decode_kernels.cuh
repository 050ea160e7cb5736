// decode_kernels.cuh -- the global kernels around block_decode.cuh: BC1 / ETC1 decode to RGBA8 and the fused
// decode + squared-error reduction.
//
// The step after the encoder in the reference harness: decompressDXT1 / decompressETC1 (Src/main.cpp:561-613) call
// DecoderBC::decodeBlockDXT1 / decodeBlockETC1 per block (Src/decoder.cpp:933-971) and getMsePsnr
// (Src/main.cpp:403-469) sums squared channel errors.  One thread owns one block, as in the encoders:
//   decode  8 bytes in (256 contiguous bytes per warp), four 16-byte row stores out (512 contiguous bytes per warp
//           and row) -- the encoder's access pattern mirrored, 4.5 B/px of HBM traffic
//   SSE     8 + 64 bytes in, nothing out: the decoded image is never written, so a multi-GiB batch is quality-checked
//           with 24 bytes of result.  CTAs walk down the image and reduce in registers -> warp -> CTA, so the whole
//           launch issues three 64-bit atomics per CTA (the first version issued three per WARP, 393 216 of them on
//           three addresses for one 8192^2 texture, and ran at 12 % of the HBM roofline because of it).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "block_decode.cuh"
#include "encode_kernels.cuh"

namespace gb {

struct DecodeParams {
    const uint8_t* blocks;
    uint8_t* rgba;          // decode target (decode kernel)
    const uint8_t* source;  // original pixels (sse kernel)
    unsigned long long* sse;  // 3 x u64 accumulators R,G,B (sse kernel)
    uint32_t bw, bh;
    uint32_t stride;        // of rgba / source, bytes
};

__device__ __forceinline__ uint2 load_block(const uint8_t* p)
{
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ void store_row(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int CODEC>
__global__ void __launch_bounds__(256) decode_kernel(const DecodeParams P)
{
    pdl_launch_dependents();   // programmatic dependent launch, as in the encoders (encode_kernels.cuh)
    pdl_wait();
    const uint32_t bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;
    if (bx >= P.bw) return;
    const uint2 blk = load_block(P.blocks + ((uint64_t)by * P.bw + bx) * 8u);
    uint32_t px[16];
    decode_block<CODEC>(blk.x, blk.y, px);
    uint8_t* o = P.rgba + (uint64_t)by * 4u * P.stride + (uint64_t)bx * 16u;
#pragma unroll
    for (int y = 0; y < 4; ++y) store_row(o + (uint64_t)y * P.stride, px[4 * y], px[4 * y + 1], px[4 * y + 2], px[4 * y + 3]);
}

// A thread may accumulate at most this many blocks in its 32-bit sums: 64 blocks x 16 pixels x 255^2 x 32 lanes
// still fits the 32-bit warp reduction (2.13e9).  The launcher sizes grid.y accordingly.
constexpr uint32_t kSseMaxBlocksPerThread = 64;

// Sum over the image of (decoded - source)^2 per channel; alpha ignored (the encoders ignore it too).
// Launched with 256 x 1 threads; blockIdx.x picks the column strip, CTAs walk block rows in steps of gridDim.y.
template <int CODEC>
__global__ void __launch_bounds__(256) block_sse_kernel(const DecodeParams P)
{
    __shared__ uint32_t partial[8][3];
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t bx = blockIdx.x * 256u + threadIdx.x;
    uint32_t sr = 0, sg = 0, sb = 0;
    if (bx < P.bw) {
#pragma unroll 1
        for (uint32_t by = blockIdx.y; by < P.bh; by += gridDim.y) {
            const uint8_t* s = P.source + (uint64_t)by * 4u * P.stride + (uint64_t)bx * 16u;
            const uint4 r0 = load_row(s);
            const uint4 r1 = load_row(s + P.stride);
            const uint4 r2 = load_row(s + 2ull * P.stride);
            const uint4 r3 = load_row(s + 3ull * P.stride);
            const uint2 blk = load_block(P.blocks + ((uint64_t)by * P.bw + bx) * 8u);
            const uint32_t src[16] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w,
                                      r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
            block_sse<CODEC>(blk.x, blk.y, src, sr, sg, sb);
        }
    }
    sr = __reduce_add_sync(0xFFFFFFFFu, sr);
    sg = __reduce_add_sync(0xFFFFFFFFu, sg);
    sb = __reduce_add_sync(0xFFFFFFFFu, sb);
    const uint32_t warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31u) == 0u) {
        partial[warp][0] = sr;
        partial[warp][1] = sg;
        partial[warp][2] = sb;
    }
    __syncthreads();
    if (threadIdx.x < 3u) {
        unsigned long long total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) total += partial[w][threadIdx.x];
        if (total) atomicAdd(&P.sse[threadIdx.x], total);
    }
}

}  // namespace gb
