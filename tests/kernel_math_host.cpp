// kernel_math_host.cpp -- TEST ONLY.  Compiles the CUDA block codec (goofy_b200/csrc/block_codec.cuh)
// for the host with the device intrinsics emulated (lanes.cuh, GOOFY_B200_HOST_EMULATION), so the
// kernel's closed-form arithmetic can be checked against the oracle on a machine without a GPU.
// This is a checker for the kernel source, not a product path: nothing in goofy_b200/ builds or loads it.
#define GOOFY_B200_HOST_EMULATION 1
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "../goofy_b200/csrc/block_codec.cuh"
#include "../goofy_b200/csrc/block_decode.cuh"

extern "C" int kernel_math_compress(int codec, unsigned char* result, const unsigned char* input, unsigned width,
                                    unsigned height, unsigned stride)
{
    const bool fused = (codec & 64) != 0;   // through encode_both (the dual-output kernel), DXT1 with the flag-byte scheme
    codec &= ~64;
    const bool viaRgb24 = (codec & 128) != 0;   // the pixels take the packed-RGB kernels' way in: 3 bytes each, widened by widen_rgb24
    codec &= ~128;
    const bool floatRef = (codec & 16) != 0;   // GOOFY_B200_FLOATREF flavours (codec 16 / 17); goofyRef accepts width % 4
    codec &= 15;
    if (width % (floatRef ? 4 : 16)) return -1;
    if (height % 4) return -2;
    uint32_t lut[256];
    for (uint32_t r = 0; r < 256; ++r) lut[r] = floatRef ? gb::etc1_control_word_ref(r) : gb::etc1_control_word(r);
    for (unsigned by = 0; by < height / 4; ++by)
        for (unsigned bx = 0; bx < width / 4; ++bx) {
            uint32_t p[16];
            for (int y = 0; y < 4; ++y) std::memcpy(&p[4 * y], input + (size_t)(4 * by + y) * stride + (size_t)bx * 16, 16);
            if (viaRgb24) {
                // what encode_rgb24_kernel sees: the block row's twelve bytes, with the bytes that FOLLOW in a packed row
                // (the next block's first pixel, or padding) spilling into the unused byte of each pixel word
                for (int y = 0; y < 4; ++y) {
                    unsigned char packed[16] = {0xA5, 0xA5, 0xA5, 0xA5, 0xA5, 0xA5, 0xA5, 0xA5, 0xA5, 0xA5, 0xA5, 0xA5, 0xA5, 0xA5, 0xA5, 0xA5};
                    for (int x = 0; x < 4; ++x) std::memcpy(packed + 3 * x, &p[4 * y + x], 3);
                    uint32_t w[3];
                    std::memcpy(w, packed, 12);
                    gb::widen_rgb24(w[0], w[1], w[2], p[4 * y], p[4 * y + 1], p[4 * y + 2], p[4 * y + 3]);
                }
            }
            uint32_t w0, w1;
            if (floatRef) {
                const gb::RefFront f = gb::analyse_ref(p, codec == 0 ? 32u : 64u);
                if (codec == 0) gb::encode_dxt1_ref(p, f, w0, w1);
                else gb::encode_etc1_ref(p, f, lut, w0, w1);
            } else {
                const gb::BlockFront f = gb::analyse(p);
                if (fused) {
                    uint32_t d0, d1, e0, e1, s0, s1;
                    gb::encode_both(p, f, lut, d0, d1, e0, e1);
                    // the single-codec entry points must agree with the fused one
                    if (codec == 0) gb::encode_dxt1<gb::kSelFlagBytes>(p, f, s0, s1);
                    else gb::encode_etc1(p, f, lut, s0, s1);
                    w0 = codec == 0 ? d0 : e0;
                    w1 = codec == 0 ? d1 : e1;
                    if (s0 != w0 || s1 != w1) return -99;
                } else if (codec == 0) gb::encode_dxt1<gb::kSelLanes>(p, f, w0, w1);
                else gb::encode_etc1(p, f, lut, w0, w1);
            }
            std::memcpy(result, &w0, 4);
            std::memcpy(result + 4, &w1, 4);
            result += 8;
        }
    return 0;
}

// Decoders (block_decode.cuh): blocks -> RGBA8 image (tight stride), codec 0 = BC1, 1 = ETC1
extern "C" int kernel_math_decode(int codec, const unsigned char* blocks, unsigned width, unsigned height, unsigned char* rgba)
{
    if (width % 4 || height % 4) return -1;
    for (unsigned by = 0; by < height / 4; ++by)
        for (unsigned bx = 0; bx < width / 4; ++bx) {
            uint32_t w[2], px[16];
            std::memcpy(w, blocks + ((size_t)by * (width / 4) + bx) * 8, 8);
            if (codec == 0) gb::decode_block<0>(w[0], w[1], px);
            else gb::decode_block<1>(w[0], w[1], px);
            for (int y = 0; y < 4; ++y) std::memcpy(rgba + ((size_t)(4 * by + y) * width + 4 * bx) * 4, &px[4 * y], 16);
        }
    return 0;
}

// Fused decode + squared error against `rgba` (tight stride): sse[0..2] = R, G, B sums
extern "C" int kernel_math_sse(int codec, const unsigned char* blocks, const unsigned char* rgba, unsigned width, unsigned height,
                               unsigned long long* sse)
{
    if (width % 4 || height % 4) return -1;
    sse[0] = sse[1] = sse[2] = 0;
    for (unsigned by = 0; by < height / 4; ++by)
        for (unsigned bx = 0; bx < width / 4; ++bx) {
            uint32_t w[2], src[16], sr = 0, sg = 0, sb = 0;
            std::memcpy(w, blocks + ((size_t)by * (width / 4) + bx) * 8, 8);
            for (int y = 0; y < 4; ++y) std::memcpy(&src[4 * y], rgba + ((size_t)(4 * by + y) * width + 4 * bx) * 4, 16);
            if (codec == 0) gb::block_sse<0>(w[0], w[1], src, sr, sg, sb);
            else gb::block_sse<1>(w[0], w[1], src, sr, sg, sb);
            sse[0] += sr; sse[1] += sg; sse[2] += sb;
        }
    return 0;
}
