// decode_kernels.cuh -- BC1 / ETC1 block decoders and the squared-error reduction on the GPU.
//
// The step after the encoder in the reference harness: decompressDXT1 / decompressETC1
// (Src/main.cpp:561-613) call DecoderBC::decodeBlockDXT1 / decodeBlockETC1 per block
// (Src/decoder.cpp:933-971 -> :819-871 for BC1, :388-678 for ETC1 differential) and getMsePsnr
// (Src/main.cpp:403-469) sums squared channel errors.  Here one thread decodes one block; a
// fused variant compares the decoded block with the source pixels without ever writing the
// decoded image, so a multi-GiB batch can be quality-checked with 12 bytes of device-to-host
// traffic.  Written from the BC1 / ETC1 format definitions, not from the reference decoder.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gb {

// 16 decoded pixels (RGBA8, alpha as the reference writes it), row-major
struct DecodedBlock {
    uint32_t px[16];
};

__device__ __forceinline__ uint32_t pack_rgba(uint32_t r, uint32_t g, uint32_t b, uint32_t a)
{
    return r | (g << 8) | (b << 16) | (a << 24);
}

// BC1: two RGB565 endpoints expanded by bit replication; c0 > c1 -> 4-colour mode with
// truncating thirds, otherwise 3 colours + transparent black (Src/decoder.cpp:798-871).
__device__ __forceinline__ DecodedBlock decode_dxt1_block(uint32_t w0, uint32_t w1)
{
    const uint32_t c0 = w0 & 0xFFFFu, c1 = w0 >> 16;
    uint32_t r[4], g[4], b[4], a[4];
    r[0] = (c0 >> 11) & 31u; g[0] = (c0 >> 5) & 63u; b[0] = c0 & 31u;
    r[1] = (c1 >> 11) & 31u; g[1] = (c1 >> 5) & 63u; b[1] = c1 & 31u;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        r[k] = (r[k] << 3) | (r[k] >> 2);
        g[k] = (g[k] << 2) | (g[k] >> 4);
        b[k] = (b[k] << 3) | (b[k] >> 2);
        a[k] = 255u;
    }
    if (c0 > c1) {
        r[2] = (2u * r[0] + r[1]) / 3u; g[2] = (2u * g[0] + g[1]) / 3u; b[2] = (2u * b[0] + b[1]) / 3u;
        r[3] = (r[0] + 2u * r[1]) / 3u; g[3] = (g[0] + 2u * g[1]) / 3u; b[3] = (b[0] + 2u * b[1]) / 3u;
        a[2] = a[3] = 255u;
    } else {
        r[2] = (r[0] + r[1]) >> 1; g[2] = (g[0] + g[1]) >> 1; b[2] = (b[0] + b[1]) >> 1;
        r[3] = g[3] = b[3] = 0u;
        a[2] = 255u; a[3] = 0u;
    }
    uint32_t pal[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) pal[k] = pack_rgba(r[k], g[k], b[k], a[k]);
    DecodedBlock d;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t k = (w1 >> (2 * i)) & 3u;
        d.px[i] = k == 0u ? pal[0] : k == 1u ? pal[1] : k == 2u ? pal[2] : pal[3];
    }
    return d;
}

__device__ __forceinline__ uint32_t clamp255(int v) { return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// ETC1 (individual and differential modes, both flip orientations).  Bytes 0-3 arrive as the
// little-endian word w0 (byte0 = R, byte1 = G, byte2 = B, byte3 = control), bytes 4-7 as w1;
// selector bit of pixel (x,y) is 4x+y in the big-endian 16-bit planes (msb bytes 4-5, lsb 6-7).
__device__ __forceinline__ DecodedBlock decode_etc1_block(uint32_t w0, uint32_t w1)
{
    const uint32_t ctl = w0 >> 24;
    const bool diff = (ctl >> 1) & 1u, flip = ctl & 1u;
    const uint32_t cw0 = (ctl >> 5) & 7u, cw1 = (ctl >> 2) & 7u;
    int base0[3], base1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t v = (w0 >> (8 * c)) & 255u;
        if (diff) {
            const uint32_t c5 = v >> 3;
            int d = (int)(v & 7u);
            d = d >= 4 ? d - 8 : d;
            const uint32_t c5b = (uint32_t)((int)c5 + d) & 31u;
            base0[c] = (int)((c5 << 3) | (c5 >> 2));
            base1[c] = (int)((c5b << 3) | (c5b >> 2));
        } else {
            base0[c] = (int)((v >> 4) * 17u);
            base1[c] = (int)((v & 15u) * 17u);
        }
    }
    const uint32_t msb = ((w1 & 0xFFu) << 8) | ((w1 >> 8) & 0xFFu);
    const uint32_t lsb = (((w1 >> 16) & 0xFFu) << 8) | (w1 >> 24);
    DecodedBlock d;
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const uint32_t bit = (uint32_t)(4 * x + y);
            const uint32_t m = (msb >> bit) & 1u, l = (lsb >> bit) & 1u;
            const bool second = flip ? (y >= 2) : (x >= 2);
            const uint32_t cw = second ? cw1 : cw0;
            // tables: small {2,5,9,13,18,24,33,47}, large {8,17,29,42,60,80,106,183}
            const uint64_t smallTab = 0x2F2118120D090502ull, largeTab = 0xB76A503C2A1D1108ull;
            const int mag = (int)(((l ? largeTab : smallTab) >> (8u * cw)) & 255u);
            const int delta = m ? -mag : mag;
            const int* base = second ? base1 : base0;
            d.px[4 * y + x] = pack_rgba(clamp255(base[0] + delta), clamp255(base[1] + delta), clamp255(base[2] + delta), 255u);
        }
    return d;
}

struct DecodeParams {
    const uint8_t* blocks;
    uint8_t* rgba;          // decode target (decode kernel)
    const uint8_t* source;  // original pixels (sse kernel)
    unsigned long long* sse;  // 3 x u64 accumulators R,G,B (sse kernel)
    uint32_t bw, bh;
    uint32_t stride;        // of rgba / source, bytes
};

template <int CODEC>
__global__ void __launch_bounds__(256) decode_kernel(const DecodeParams P)
{
    const uint32_t bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;
    if (bx >= P.bw) return;
    const uint2 blk = *reinterpret_cast<const uint2*>(P.blocks + ((uint64_t)by * P.bw + bx) * 8u);
    const DecodedBlock d = CODEC == 0 ? decode_dxt1_block(blk.x, blk.y) : decode_etc1_block(blk.x, blk.y);
    uint8_t* o = P.rgba + (uint64_t)by * 4u * P.stride + (uint64_t)bx * 16u;
#pragma unroll
    for (int y = 0; y < 4; ++y)
        *reinterpret_cast<uint4*>(o + (uint64_t)y * P.stride) = make_uint4(d.px[4 * y], d.px[4 * y + 1], d.px[4 * y + 2], d.px[4 * y + 3]);
}

// Sum over the image of (decoded - source)^2 per channel; alpha ignored (the encoders ignore it too).
template <int CODEC>
__global__ void __launch_bounds__(256) block_sse_kernel(const DecodeParams P)
{
    const uint32_t bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;
    uint32_t sr = 0, sg = 0, sb = 0;
    if (bx < P.bw) {
        const uint2 blk = *reinterpret_cast<const uint2*>(P.blocks + ((uint64_t)by * P.bw + bx) * 8u);
        const DecodedBlock d = CODEC == 0 ? decode_dxt1_block(blk.x, blk.y) : decode_etc1_block(blk.x, blk.y);
        const uint8_t* s = P.source + (uint64_t)by * 4u * P.stride + (uint64_t)bx * 16u;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const uint4 row = *reinterpret_cast<const uint4*>(s + (uint64_t)y * P.stride);
            const uint32_t src[4] = {row.x, row.y, row.z, row.w};
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const uint32_t ad = __vabsdiffu4(d.px[4 * y + x], src[x]);  // |difference| per byte
                sr = __dp4a(ad, ad & 0x000000FFu, sr);
                sg = __dp4a(ad, ad & 0x0000FF00u, sg);
                sb = __dp4a(ad, ad & 0x00FF0000u, sb);
            }
        }
    }
    sr = __reduce_add_sync(0xFFFFFFFFu, sr);
    sg = __reduce_add_sync(0xFFFFFFFFu, sg);
    sb = __reduce_add_sync(0xFFFFFFFFu, sb);
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(&P.sse[0], (unsigned long long)sr);
        atomicAdd(&P.sse[1], (unsigned long long)sg);
        atomicAdd(&P.sse[2], (unsigned long long)sb);
    }
}

}  // namespace gb
