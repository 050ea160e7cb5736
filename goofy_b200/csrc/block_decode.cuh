// block_decode.cuh -- BC1 / ETC1 block decoding as byte-permute table lookups, one thread per block.
//
// The step after the encoders in the reference harness: DecoderBC::decodeBlockDXT1 / decodeBlockETC1
// (Src/decoder.cpp:933-971 -> :798-871 BC1 colour block, :388-678 ETC1) and the squared-error sums of
// getMsePsnr (Src/main.cpp:403-469).  Written from the BC1 / ETC1 format definitions, not from that decoder.
//
// Both formats are "a small palette + a selector per pixel".  The palette is kept PLANAR -- one register per
// channel holding that channel of the four palette colours (ETC1: one such register per sub-block) -- so a
// single PRMT whose selector nibbles are the pixel selectors fetches one channel of FOUR pixels:
//   decode   row of 4 pixels = 3-4 lookups + 8 permutes to interleave the planar channels back into RGBA words
//   SSE      the source row is de-interleaved instead (7 permutes); |decoded - source| is one VABSDIFF4 and the
//            sum of squares one IDP.4A per channel per row -- the decoded image is never formed
// Everything is built from lanes.cuh primitives, so tests/kernel_math_host.cpp checks it without a GPU.
#pragma once
#include "lanes.cuh"

namespace gb {

// Planar palette: channel c of palette entry k is byte k of ch[c][0] (entries 0..3) / ch[c][1] (entries 4..7).
// BC1 uses entries 0..3 only (both registers equal); ETC1 entries 4s + (2*msb + lsb) for sub-block s.
struct PlanarPalette {
    uint32_t r[2], g[2], b[2], a[2];
};

// floor(x / 3) for 0 <= x <= 765 (exhaustively checked in tests/test_closed_forms.py)
GB_DEV uint32_t third(uint32_t x) { return (x * 43691u) >> 17; }

// BC1 colour block (Src/decoder.cpp:798-871): RGB565 endpoints expanded by bit replication; c0 > c1 selects
// the 4-colour mode with truncating thirds, otherwise entry 2 is the truncated mean and entry 3 transparent black.
GB_DEV PlanarPalette bc1_palette(uint32_t w0)
{
    const uint32_t c0 = w0 & 0xFFFFu, c1 = w0 >> 16;
    const uint32_t r0 = ((c0 >> 8) & 0xF8u) | (c0 >> 13), r1 = ((c1 >> 8) & 0xF8u) | (c1 >> 13);
    const uint32_t g0 = ((c0 >> 3) & 0xFCu) | ((c0 >> 9) & 3u), g1 = ((c1 >> 3) & 0xFCu) | ((c1 >> 9) & 3u);
    const uint32_t b0 = ((c0 << 3) & 0xF8u) | ((c0 >> 2) & 7u), b1 = ((c1 << 3) & 0xF8u) | ((c1 >> 2) & 7u);
    const bool four = c0 > c1;
    const uint32_t r2 = four ? third(2u * r0 + r1) : (r0 + r1) >> 1, r3 = four ? third(r0 + 2u * r1) : 0u;
    const uint32_t g2 = four ? third(2u * g0 + g1) : (g0 + g1) >> 1, g3 = four ? third(g0 + 2u * g1) : 0u;
    const uint32_t b2 = four ? third(2u * b0 + b1) : (b0 + b1) >> 1, b3 = four ? third(b0 + 2u * b1) : 0u;
    PlanarPalette p;
    p.r[0] = p.r[1] = r0 | (r1 << 8) | (r2 << 16) | (r3 << 24);
    p.g[0] = p.g[1] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    p.b[0] = p.b[1] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    p.a[0] = p.a[1] = four ? 0xFFFFFFFFu : 0x00FFFFFFu;
    return p;
}

// Row selectors of a BC1 block: 2 bits per pixel, row major -> one nibble per pixel (16 bits per row).
GB_DEV void bc1_row_selectors(uint32_t w1, uint32_t (&sel)[4])
{
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t x = prmt(w1, 0u, half ? 0x4342u : 0x4140u);   // bytes (row 2h, 0, row 2h+1, 0)
        x = (x | (x << 4)) & 0x0F0F0F0Fu;
        x = (x | (x << 2)) & 0x33333333u;
        sel[2 * half] = x & 0xFFFFu;
        sel[2 * half + 1] = x >> 16;
    }
}

// ETC1 (individual and differential modes, both flip orientations).  w0 = bytes (R, G, B, control).
// Entry 4s + k of the palette is sub-block s with modifier k = 2*msb + lsb: +small, +large, -small, -large.
GB_DEV PlanarPalette etc1_palette(uint32_t w0)
{
    const uint32_t ctl = w0 >> 24;
    // base colours of the two sub-blocks, three channels at once in the low three bytes (no byte ever carries)
    const uint32_t c5 = (w0 >> 3) & 0x1F1F1Fu;
    // differential: second = (first + signed 3-bit delta) mod 32; -8 == +24 (mod 32), so a set sign bit adds 28 - 4
    const uint32_t c5b = (c5 + (w0 & 0x030303u) + ((w0 & 0x040404u) * 7u)) & 0x1F1F1Fu;
    const uint32_t d0 = (c5 << 3) | ((c5 >> 2) & 0x070707u), d1 = (c5b << 3) | ((c5b >> 2) & 0x070707u);
    const uint32_t i0 = ((w0 >> 4) & 0x0F0F0Fu) * 17u, i1 = (w0 & 0x0F0F0Fu) * 17u;
    const bool diff = (ctl & 2u) != 0u;
    const uint32_t base[2] = {diff ? d0 : i0, diff ? d1 : i1};
    // modifier tables {2,5,9,13,18,24,33,47} / {8,17,29,42,60,80,106,183}, fetched by one byte permute each
    const uint32_t cw[2] = {(ctl >> 5) & 7u, (ctl >> 2) & 7u};
    PlanarPalette p;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const uint32_t small = prmt(0x0D090502u, 0x2F211812u, cw[s]) & 0xFFu;
        const uint32_t large = prmt(0x2A1D1108u, 0xB76A503Cu, cw[s]) & 0xFFu;
        const uint32_t dpos = small | (large << 16);     // lanes (+small, +large)
        const uint32_t dneg = 0x10000u - dpos;           // lanes (-small, -large): small > 0 borrows one from the high lane
        uint32_t planar[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t v = prmt(base[s], 0u, 0x4040u | (uint32_t)c | ((uint32_t)c << 8));   // lanes (v, v)
            const uint32_t pos = addclamp_s16x2(v, dpos, 0x00FF00FFu);
            const uint32_t neg = addclamp_s16x2(v, dneg, 0x00FF00FFu);
            planar[c] = prmt(pos, neg, 0x6420u);         // bytes (v+small, v+large, v-small, v-large), clamped
        }
        p.r[s] = planar[0];
        p.g[s] = planar[1];
        p.b[s] = planar[2];
        p.a[s] = 0xFFFFFFFFu;
    }
    return p;
}

// Row selectors of an ETC1 block: nibble of pixel x in row y = 4*subblock + 2*msb + lsb; the selector bit of
// pixel (x,y) is bit 4x+y of the big-endian 16-bit planes (msb plane bytes 4-5, lsb plane bytes 6-7).
GB_DEV void etc1_row_selectors(uint32_t w0, uint32_t w1, uint32_t (&sel)[4])
{
    const uint32_t planes = prmt(w1, 0u, 0x0123u);   // lsb plane | msb plane << 16
    const bool flip = ((w0 >> 24) & 1u) != 0u;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
        const uint32_t sub = flip ? (y >= 2 ? 0x4444u : 0u) : 0x4400u;
        sel[y] = ((planes >> y) & 0x1111u) + ((planes >> (15 + y)) & 0x2222u) + sub;
    }
}

// One channel of the four pixels of a row
GB_DEV uint32_t lookup4(const uint32_t (&ch)[2], uint32_t sel) { return prmt(ch[0], ch[1], sel); }

// planar (R0..R3), (G0..G3), (B0..B3), (A0..A3) -> four RGBA words
GB_DEV void interleave4(uint32_t r4, uint32_t g4, uint32_t b4, uint32_t a4, uint32_t (&px)[4])
{
    const uint32_t rgLo = prmt(r4, g4, 0x5140u), rgHi = prmt(r4, g4, 0x7362u);   // (R0,G0,R1,G1), (R2,G2,R3,G3)
    const uint32_t baLo = prmt(b4, a4, 0x5140u), baHi = prmt(b4, a4, 0x7362u);
    px[0] = prmt(rgLo, baLo, 0x5410u);
    px[1] = prmt(rgLo, baLo, 0x7632u);
    px[2] = prmt(rgHi, baHi, 0x5410u);
    px[3] = prmt(rgHi, baHi, 0x7632u);
}

// Adds the squared channel errors of one row (decoded planar r4/g4/b4 against four source RGBA words).
GB_DEV void row_sse(uint32_t r4, uint32_t g4, uint32_t b4, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3,
                    uint32_t& sr, uint32_t& sg, uint32_t& sb)
{
    const uint32_t rg01 = prmt(s0, s1, 0x5410u), ba01 = prmt(s0, s1, 0x7632u);
    const uint32_t rg23 = prmt(s2, s3, 0x5410u), ba23 = prmt(s2, s3, 0x7632u);
    const uint32_t dr = absdiff_u8x4(r4, prmt(rg01, rg23, 0x6420u));
    const uint32_t dg = absdiff_u8x4(g4, prmt(rg01, rg23, 0x7531u));
    const uint32_t db = absdiff_u8x4(b4, prmt(ba01, ba23, 0x6420u));
    sr = dp4a(dr, dr, sr);
    sg = dp4a(dg, dg, sg);
    sb = dp4a(db, db, sb);
}

// CODEC 0 = BC1, 1 = ETC1
template <int CODEC>
GB_DEV void decode_block(uint32_t w0, uint32_t w1, uint32_t (&px)[16])
{
    const PlanarPalette pal = CODEC == 0 ? bc1_palette(w0) : etc1_palette(w0);
    uint32_t sel[4];
    if (CODEC == 0) bc1_row_selectors(w1, sel);
    else etc1_row_selectors(w0, w1, sel);
#pragma unroll
    for (int y = 0; y < 4; ++y) {
        uint32_t row[4];
        interleave4(lookup4(pal.r, sel[y]), lookup4(pal.g, sel[y]), lookup4(pal.b, sel[y]),
                    CODEC == 0 ? lookup4(pal.a, sel[y]) : 0xFFFFFFFFu, row);
#pragma unroll
        for (int x = 0; x < 4; ++x) px[4 * y + x] = row[x];
    }
}

// Sum of squared (decoded - source) per channel over one block; src = the 16 source pixels, row major.
template <int CODEC>
GB_DEV void block_sse(uint32_t w0, uint32_t w1, const uint32_t (&src)[16], uint32_t& sr, uint32_t& sg, uint32_t& sb)
{
    const PlanarPalette pal = CODEC == 0 ? bc1_palette(w0) : etc1_palette(w0);
    uint32_t sel[4];
    if (CODEC == 0) bc1_row_selectors(w1, sel);
    else etc1_row_selectors(w0, w1, sel);
#pragma unroll
    for (int y = 0; y < 4; ++y)
        row_sse(lookup4(pal.r, sel[y]), lookup4(pal.g, sel[y]), lookup4(pal.b, sel[y]), src[4 * y], src[4 * y + 1],
                src[4 * y + 2], src[4 * y + 3], sr, sg, sb);
}

}  // namespace gb
