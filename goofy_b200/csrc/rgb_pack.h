// rgb_pack.h -- host-side helper of the drop-in host path (host_pipeline.cuh): RGBA8 rows -> packed RGB8 rows.
//
// The encoders ignore alpha (GoofyTC/goofy_tc.h:297 drops it; the reference's own loader forces it to 0xFF,
// Src/main.cpp:328-335), yet a host image carries it over PCIe: 4 of the 4.5 bytes per pixel that cross the link.  The
// host path is bound by exactly that link, so the staging copy it has to make anyway (pageable buffers), or a staging
// copy it adds (pinned buffers, when host threads are free), drops the alpha byte on the way: 3 B/px cross instead of 4
// and the rgb24 kernels (encode_kernels.cuh: encode_rgb24_kernel) expand the pixels again in registers.
// This is data marshalling, not encoding: there is no CPU encoder anywhere in this library.
//
// Plain C++ (no CUDA), so tests/rgb_pack_host.cpp can check it on a machine without a GPU.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define GB_PACK_X86 1
#endif

namespace gbpack {

// Portable form: four pixels (16 bytes) -> three 32-bit words, little endian.
inline void pack_row_scalar(uint8_t* dst, const uint8_t* src, size_t pixels)
{
    for (size_t i = 0; i + 4u <= pixels; i += 4u) {
        uint32_t p[4];
        std::memcpy(p, src + 4u * i, 16);
        const uint32_t w[3] = {(p[0] & 0xFFFFFFu) | (p[1] << 24), ((p[1] >> 8) & 0xFFFFu) | (p[2] << 16), ((p[2] >> 16) & 0xFFu) | (p[3] << 8)};
        std::memcpy(dst + 3u * i, w, 12);
    }
}

#ifdef GB_PACK_X86
// SSSE3: 16 pixels (64 bytes) -> 48 bytes with six byte shuffles and three ORs.  Every output vector takes its bytes
// from two neighbouring input vectors; a shuffle-control byte with the top bit set writes zero.
// (Software prefetch 512 / 1024 bytes ahead of the loads was measured on the GPU box: no consistent gain, session P.)
__attribute__((target("ssse3"))) inline void pack_row_ssse3(uint8_t* dst, const uint8_t* src, size_t pixels, bool streaming)
{
    const __m128i a0 = _mm_setr_epi8(0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1);     // a -> out0[0..11]
    const __m128i b0 = _mm_setr_epi8(-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 1, 2, 4);  // b -> out0[12..15]
    const __m128i b1 = _mm_setr_epi8(5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1, -1, -1, -1, -1);  // b -> out1[0..7]
    const __m128i c1 = _mm_setr_epi8(-1, -1, -1, -1, -1, -1, -1, -1, 0, 1, 2, 4, 5, 6, 8, 9);      // c -> out1[8..15]
    const __m128i c2 = _mm_setr_epi8(10, 12, 13, 14, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);  // c -> out2[0..3]
    const __m128i d2 = _mm_setr_epi8(-1, -1, -1, -1, 0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14);      // d -> out2[4..15]
    const bool alignedOut = streaming && ((uintptr_t)dst & 15u) == 0u;
    size_t i = 0;
    for (; i + 16u <= pixels; i += 16u) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + 4u * i));
        const __m128i b = _mm_loadu_si128((const __m128i*)(src + 4u * i + 16u));
        const __m128i c = _mm_loadu_si128((const __m128i*)(src + 4u * i + 32u));
        const __m128i d = _mm_loadu_si128((const __m128i*)(src + 4u * i + 48u));
        const __m128i o0 = _mm_or_si128(_mm_shuffle_epi8(a, a0), _mm_shuffle_epi8(b, b0));
        const __m128i o1 = _mm_or_si128(_mm_shuffle_epi8(b, b1), _mm_shuffle_epi8(c, c1));
        const __m128i o2 = _mm_or_si128(_mm_shuffle_epi8(c, c2), _mm_shuffle_epi8(d, d2));
        __m128i* o = (__m128i*)(dst + 3u * i);
        if (alignedOut) {
            _mm_stream_si128(o, o0);
            _mm_stream_si128(o + 1, o1);
            _mm_stream_si128(o + 2, o2);
        } else {
            _mm_storeu_si128(o, o0);
            _mm_storeu_si128(o + 1, o1);
            _mm_storeu_si128(o + 2, o2);
        }
    }
    if (i < pixels) pack_row_scalar(dst + 3u * i, src + 4u * i, pixels - i);   // widths that are multiples of 4 only
}

inline bool have_ssse3()
{
    static const bool yes = __builtin_cpu_supports("ssse3") != 0;
    return yes;
}
#endif

// `pixels` must be a multiple of 4 (the encoders' width contract is 16, the float-reference flavour's 4); dst and src
// rows must not overlap.
inline void pack_row(uint8_t* dst, const uint8_t* src, size_t pixels, bool streaming)
{
#ifdef GB_PACK_X86
    if (have_ssse3()) {
        pack_row_ssse3(dst, src, pixels, streaming);
        return;
    }
#endif
    (void)streaming;
    pack_row_scalar(dst, src, pixels);
}

// rows [r0, r1) of an image: source rows `srcPitch` bytes apart, packed rows `dstPitch` bytes apart
inline void pack_rows(uint8_t* dst, size_t dstPitch, const uint8_t* src, size_t srcPitch, size_t pixels, size_t r0, size_t r1, bool streaming)
{
    for (size_t r = r0; r < r1; ++r) pack_row(dst + r * dstPitch, src + r * srcPitch, pixels, streaming);
#ifdef GB_PACK_X86
    if (streaming) _mm_sfence();   // the DMA engine reads these rows next
#endif
}

}  // namespace gbpack
