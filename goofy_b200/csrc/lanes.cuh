// lanes.cuh -- the handful of packed-integer primitives the block encoders are built from.
//
// On the device every one of these is a single sm_100a instruction (checked with
// cuobjdump -sass, see DESIGN.md "instruction budget"):
//   dp4a        IDP.4A.U8.U8      4 x (u8*u8) + u32
//   prmt        PRMT              byte permute of two words
//   min3/max3   VIMNMX3.U16x2     3-input min/max on two u16 lanes
//   min2/max2   VIMNMX.U16x2
//   bitsel      LOP3.LUT          (a & m) | (b & ~m)
//   sign_bytes  PRMT              byte permute in sign-replicate mode: four flag bytes (0x00 / 0xFF) from four sign bits
//   dp4a_su     IDP.4A.S8.U8      4 x (s8*u8) + s32: subtracts the weights of the set flag bytes
//   absdiff_u8x4 VABSDIFF4.U8     |a - b| on four unsigned bytes
//   dp2a_lo/hi  IDP.2A.LO/HI      2 x (u16*u8) + u32
// These stand in for the SSE2 pminub/pmaxub/pavgb/punpck*/pmovmskb sequences of the
// reference (GoofyTC/goofy_tc.h:170-396); the encoders do NOT transliterate those ops,
// they use closed forms on u16x2 lanes.
//
// GOOFY_B200_HOST_EMULATION is defined ONLY by tests/kernel_math_host.cpp, which compiles
// block_codec.cuh with g++ to check the kernel arithmetic against the oracle on a machine
// without a GPU.  The product library is always built by nvcc and never takes that branch.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GB_DEV __device__ __forceinline__
#else
#ifndef GOOFY_B200_HOST_EMULATION
#error "lanes.cuh: device code only (host emulation exists for tests/kernel_math_host.cpp alone)"
#endif
#define GB_DEV static inline
#endif

namespace gb {

#if defined(__CUDACC__)

GB_DEV uint32_t dp4a(uint32_t a, uint32_t w, uint32_t c) { return __dp4a(a, w, c); }
GB_DEV uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
GB_DEV uint32_t min3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
GB_DEV uint32_t max3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
GB_DEV uint32_t min2_u16x2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
GB_DEV uint32_t max2_u16x2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
// per-lane clamp(a + b, 0, c) on two signed 16-bit lanes
GB_DEV uint32_t addclamp_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2_relu(a, b, c); }
// (a ^ b) & m as ONE LOP3.  Written in PTX so the compiler cannot re-associate the mask past a
// following shift (it would, and that costs an extra instruction per average).
GB_DEV uint32_t xor_and(uint32_t a, uint32_t b, uint32_t m)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(r) : "r"(a), "r"(b), "r"(m));
    return r;
}
// ~(a | b) as ONE LOP3 (the compiler otherwise emits OR then NOT in front of a LEA.HI)
GB_DEV uint32_t nor(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, 0, 0x03;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// per-lane clamp(a + b, 0, c) on one signed 32-bit value
GB_DEV int addclamp_s32(int a, int b, int c) { return __viaddmin_s32_relu(a, b, c); }
// |a - b| per unsigned byte
GB_DEV uint32_t absdiff_u8x4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }
// IDP.2A: c + w.lo16 * (byte 0 of px) + w.hi16 * (byte 1 of px)   resp. bytes 2 and 3
GB_DEV uint32_t dp2a_lo(uint32_t w, uint32_t px, uint32_t c) { return __dp2a_lo(w, px, c); }
GB_DEV uint32_t dp2a_hi(uint32_t w, uint32_t px, uint32_t c) { return __dp2a_hi(w, px, c); }
// Byte permute whose selector nibbles have bit 3 set: result byte = the MSB of the chosen byte of (a, b),
// replicated over all 8 bits (0x00 or 0xFF).  kSel is the usual 4-nibble selector with 8 added to each nibble.
template <uint32_t kSel>
GB_DEV uint32_t sign_bytes(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "n"(kSel));
    return r;
}
// c + sum of (signed byte of a) * (unsigned byte of w): with flag bytes 0x00 / 0xFF (= 0 / -1) this is
// c minus the weights of the set flags.
GB_DEV uint32_t dp4a_su(uint32_t a, uint32_t w, uint32_t c)
{
    int r;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(w), "r"((int)c));
    return (uint32_t)r;
}

#else  // host emulation (tests only)

GB_DEV uint32_t dp4a(uint32_t a, uint32_t w, uint32_t c)
{
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 255u) * ((w >> (8 * i)) & 255u);
    return c;
}
GB_DEV uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 255u) << (8 * i);
    return r;
}
GB_DEV uint32_t lanes_(uint32_t lo, uint32_t hi) { return (lo & 0xFFFFu) | (hi << 16); }
GB_DEV uint32_t min2_u16x2(uint32_t a, uint32_t b)
{
    uint32_t al = a & 0xFFFFu, bl = b & 0xFFFFu, ah = a >> 16, bh = b >> 16;
    return lanes_(al < bl ? al : bl, ah < bh ? ah : bh);
}
GB_DEV uint32_t max2_u16x2(uint32_t a, uint32_t b)
{
    uint32_t al = a & 0xFFFFu, bl = b & 0xFFFFu, ah = a >> 16, bh = b >> 16;
    return lanes_(al > bl ? al : bl, ah > bh ? ah : bh);
}
GB_DEV uint32_t min3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return min2_u16x2(min2_u16x2(a, b), c); }
GB_DEV uint32_t max3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return max2_u16x2(max2_u16x2(a, b), c); }
GB_DEV uint32_t addclamp_s16x2(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r = 0;
    for (int i = 0; i < 2; ++i) {
        int v = (int)(int16_t)(a >> (16 * i)) + (int)(int16_t)(b >> (16 * i));
        int hi = (int)(int16_t)(c >> (16 * i));
        if (v > hi) v = hi;
        if (v < 0) v = 0;
        r |= ((uint32_t)v & 0xFFFFu) << (16 * i);
    }
    return r;
}

GB_DEV uint32_t xor_and(uint32_t a, uint32_t b, uint32_t m) { return (a ^ b) & m; }
GB_DEV uint32_t nor(uint32_t a, uint32_t b) { return ~(a | b); }
GB_DEV int addclamp_s32(int a, int b, int c)
{
    int v = a + b;
    v = v > c ? c : v;
    return v < 0 ? 0 : v;
}
GB_DEV uint32_t dp2a_lo(uint32_t w, uint32_t px, uint32_t c)
{
    return c + (w & 0xFFFFu) * (px & 255u) + (w >> 16) * ((px >> 8) & 255u);
}
GB_DEV uint32_t dp2a_hi(uint32_t w, uint32_t px, uint32_t c)
{
    return c + (w & 0xFFFFu) * ((px >> 16) & 255u) + (w >> 16) * (px >> 24);
}
GB_DEV uint32_t absdiff_u8x4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const int x = (int)((a >> (8 * i)) & 255u), y = (int)((b >> (8 * i)) & 255u);
        r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i);
    }
    return r;
}
template <uint32_t kSel>
GB_DEV uint32_t sign_bytes(uint32_t a, uint32_t b)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t s = (kSel >> (4 * i)) & 15u;
        uint32_t byte = (uint32_t)(v >> (8 * (s & 7u))) & 255u;
        if (s & 8u) byte = (byte & 0x80u) ? 0xFFu : 0u;
        r |= byte << (8 * i);
    }
    return r;
}
GB_DEV uint32_t dp4a_su(uint32_t a, uint32_t w, uint32_t c)
{
    for (int i = 0; i < 4; ++i) c += (uint32_t)((int)(int8_t)(a >> (8 * i)) * (int)((w >> (8 * i)) & 255u));
    return c;
}

#endif

// (a & m) | (b & ~m): one LOP3
GB_DEV uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m) { return (a & m) | (b & ~m); }

}  // namespace gb
