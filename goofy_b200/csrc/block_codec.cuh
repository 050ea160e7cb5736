// block_codec.cuh -- one thread encodes one 4x4 block held in 16 registers.
//
// What is computed is exactly what the reference's SSE2 tile kernel computes per block
// (GoofyTC/goofy_tc.h:1069-1494); how it is computed is different.  The reference spends
// most of its instructions on byte unpack/transposes that exist only because it processes a
// 16x4 tile across 16 byte lanes.  Here each thread owns a whole block, so the pipeline is
// re-derived on two 16-bit lanes per register with closed forms whose equivalence to the
// reference's rounded cascades is proven exhaustively in tests/test_closed_forms.py:
//
//   Y(R,G,B) = avg(avg(R,B),G)              == (R + 2G + B + 3) >> 2        (goofy_tc.h:1166,1205)
//   to5(v)   = avg(avg(avg(sat-(v,8),0),0),0) == (max(v,1) - 1) >> 3        (goofy_tc.h:1309-1311,1462)
//   qt       = sat+(quarter, eighth)        == ((r+3)>>2) + ((r+7)>>3)      (goofy_tc.h:1184-1190)
//   avg(a,b) = (a+b+1)>>1                   == ~floor_avg(~a,~b), per byte  (goofy_tc.h:645-654)
//
// Pixel classification never shifts Y down: with S = R+2G+B+3 (one IDP.4A per pixel) and
// e = S - 4*mid, the reference's masks (goofy_tc.h:1212-1224) become
//   Gez  <=> e >= 0                     Lqt  <=> 4 - 4qt <= e < 4qt
// Two pixels share a register as biased u16 lanes (lane = e + 0x4000); each test is then a
// packed add whose carry lands in bit 14/15 of the lane.
#pragma once
#include "lanes.cuh"

namespace gb {

constexpr uint32_t kLuma = 0x00010201u;   // dp4a weights for bytes (R,G,B,A): R + 2G + B

// min / max of 16 words, two u16 lanes each, in 8 three-input ops
template <bool kMax>
GB_DEV uint32_t reduce16(const uint32_t (&v)[16])
{
    auto m3 = [](uint32_t a, uint32_t b, uint32_t c) { return kMax ? max3_u16x2(a, b, c) : min3_u16x2(a, b, c); };
    const uint32_t t0 = m3(v[0], v[1], v[2]);
    const uint32_t t1 = m3(v[3], v[4], v[5]);
    const uint32_t t2 = m3(v[6], v[7], v[8]);
    const uint32_t t3 = m3(v[9], v[10], v[11]);
    const uint32_t t4 = m3(v[12], v[13], v[14]);
    const uint32_t t5 = m3(t0, t1, v[15]);
    const uint32_t t6 = m3(t2, t3, t4);
    return kMax ? max2_u16x2(t5, t6) : min2_u16x2(t5, t6);
}

// Packed RGB8 (3 bytes per pixel) -> the pixel words the codecs take: four pixels of one block row arrive as the three
// words w0 = R0 G0 B0 R1 | w1 = G1 B1 R2 G2 | w2 = B2 R3 G3 B3.  Byte 3 of every result is whatever followed the pixel;
// nothing below reads it (the dp4a weights of that byte are zero, the min / max lanes that carry it are never extracted).
// Two byte permutes and a shift; tests/kernel_math_host.cpp runs every fixture through it on the host.
GB_DEV void widen_rgb24(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t& p0, uint32_t& p1, uint32_t& p2, uint32_t& p3)
{
    p0 = w0;
    p1 = prmt(w0, w1, 0x6543u);
    p2 = prmt(w1, w2, 0x5432u);
    p3 = w2 >> 8;
}

// Everything the two codecs share: bounding box and the per-block brightness scalars.
struct BlockFront {
    // A u16 lane compare orders by its HIGH byte first, so the raw pixel word (lanes G:R | A:B)
    // yields min/max G in byte 1, and the word shifted left by 8 (lanes R:0 | B:G) yields
    // min/max R in byte 1 and min/max B in byte 3.  No per-channel unpack is needed.
    uint32_t mnG, mxG;    // byte1 = min / max G
    uint32_t mnRB, mxRB;  // byte1 = min / max R, byte3 = min / max B, byte0 = 0
    uint32_t range;       // max(maxY - minY, 8)                  goofy_tc.h:1176
    uint32_t mid;         // avg(minY, maxY)                      goofy_tc.h:1179
    uint32_t laneBias;    // dp4a accumulator that makes lane = (R+2G+B+3) - 4*mid + 0x4000
    uint32_t kLo, kHi;    // packed add constants: bit15 of (lane + kLo) <=> e >= 4 - 4qt, of (lane + kHi) <=> e >= 4qt
    // flag-byte scheme: lane = e + 0x8000 (bias fbG), whose bit 15 is G (e >= 0) as it stands
    uint32_t fbG;         // lanes_of accumulator: R+2G+B + fbG = e + 0x8000 (one lane's worth)
    uint32_t fbB;         // lane - fbB : bit15 <=> e >= 4qt       (both lanes)
    uint32_t fbNa;        // fbNa - lane: bit15 <=> e < 4 - 4qt    (both lanes)
};

GB_DEV BlockFront analyse(const uint32_t (&p)[16])
{
    BlockFront f;
    uint32_t q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) q[i] = p[i] << 8;
    f.mnG = reduce16<false>(p);
    f.mxG = reduce16<true>(p);
    f.mnRB = reduce16<false>(q);
    f.mxRB = reduce16<true>(q);

    // brightness of the two box corners (goofy_tc.h:1151-1171)
    const uint32_t mnPix = prmt(f.mnRB, f.mnG, 0x3351);  // bytes (minR, minG, minB, .)
    const uint32_t mxPix = prmt(f.mxRB, f.mxG, 0x3351);
    const uint32_t minY = dp4a(mnPix, kLuma, 3u) >> 2;
    const uint32_t maxY = dp4a(mxPix, kLuma, 3u) >> 2;

    const uint32_t spread = maxY - minY;
    f.range = spread > 8u ? spread : 8u;
    f.mid = (minY + maxY + 1u) >> 1;
    const uint32_t qt = ((f.range + 3u) >> 2) + ((f.range + 7u) >> 3);  // 3..96, goofy_tc.h:1184-1190

    const uint32_t q4 = qt << 2;
    f.laneBias = 0x4003u - (f.mid << 2);
    f.kLo = (0x3FFCu + q4) * 0x10001u;
    f.kHi = (0x4000u - q4) * 0x10001u;
    // e = R+2G+B + 3 - 4 mid; every biased lane stays inside 1..0xFFFE, so the two lanes never exchange a carry
    f.fbG = 0x8003u - (f.mid << 2);
    f.fbB = q4 * 0x10001u;
    f.fbNa = (0x10003u - q4) * 0x10001u;   // (0x8000 + 3 - 4qt) - e, per lane
    return f;
}

// Two pixels -> one register of biased brightness lanes (low lane = a, high lane = b).
GB_DEV uint32_t lanes_of(uint32_t a, uint32_t b, uint32_t laneBias)
{
    const uint32_t hi = dp4a(b, kLuma, laneBias);
    return dp4a(a, kLuma, hi * 65536u + laneBias);
}

// bit15 of each lane <=> Lqt (|Y - mid| < qt)
GB_DEV uint32_t lqt_lanes(uint32_t e, const BlockFront& f) { return (e + f.kLo) ^ (e + f.kHi); }

// How the per-pixel flags (Gez, Lqt) are gathered into the output words.  Both give the same bits.
//   kSelLanes      two pixels per register as biased u16 lanes, shift-and-insert accumulators on the integer ALU
//                  pipe (DXT1 indices only; what the DRAM-bound DXT1 kernel uses)
//   kSelFlagBytes  lanes as above, but the flags leave the lanes as 0x00 / 0xFF BYTES (one sign-replicating
//                  PRMT per four flags) and are weighted into place by IDP.4A on the multiply pipe: 12 integer-ALU
//                  instructions per block instead of 32-50, and the same twelve flag words feed both codecs
//                  (ETC1s, dual-output and float-reference kernels).
// Two earlier ETC1s plane gatherers -- lanes with shift-and-insert accumulators, and one pixel at a time on the
// multiply pipe with a funnel shift per flag -- were 16-20 % longer and are in the git history.
enum Selectors : int { kSelLanes = 0, kSelFlagBytes = 2 };

// Flag-byte scheme.  Per pixel three threshold flags, each the sign bit of a u16 lane (e = S - 4*mid):
//   G  = e >= 0        NA = e < 4 - 4qt        B = e >= 4qt          (NA and B exclude each other; B implies G)
// so that   !Gez = 1 - G,   Lqt = 1 - NA - B,   !Lqt = NA + B.
// Row y is walked as two pairs, (x=0,1) and (x=2,3); per row three flag words:
//   outerL / outerR = bytes (B_lo, B_hi, NA_lo, NA_hi) of the left / right pair,   gez = bytes (G_0, G_1, G_2, G_3).
// dp4a_su subtracts the weight of every set flag, hence the accumulators count DOWN from all-ones:
//   DXT1 index byte of row y = 255 - sum_x 4^x (2 NA + 2 B + G)                      (2*Lqt + !Gez per pixel)
//   ETC1 planes, pixel (x,y) at bit ((x^2)<<2)+y: x in {2,3} fills plane byte 0, x in {0,1} plane byte 1, with
//   weights 2^y and 2^(4+y) inside the byte; the !Gez plane counts down from 0xFF per byte, the !Lqt plane up.
// The IDP weights live in constant memory on the device: the compiler then keeps them in uniform registers
// across the row loop of the kernels instead of re-creating each one (a UMOV apiece) for every block.
#if defined(__CUDACC__)
#define GB_WEIGHT_TABLE __constant__
#else
#define GB_WEIGHT_TABLE static const
#endif
GB_WEIGHT_TABLE uint32_t kFlagWeights[17] = {
    0x40100401u, 0x08020802u, 0x80208020u,                                  // DXT1: gez, outerL, outerR
    0x10011001u << 0, 0x10011001u << 1, 0x10011001u << 2, 0x10011001u << 3,  // ETC1 !Lqt plane, row y
    0x00001001u << 0, 0x00001001u << 1, 0x00001001u << 2, 0x00001001u << 3,  // ETC1 !Gez plane byte 1 (x = 0,1), row y
    0x10010000u << 0, 0x10010000u << 1, 0x10010000u << 2, 0x10010000u << 3,  // ETC1 !Gez plane byte 0 (x = 2,3), row y
    0x10012002u, 0x10012002u << 2,   // ETC1-only kernel: !Gez flags of one pair from rows (y+1, y), y = 0 and y = 2
};

// kWeights = dp4a weights of the brightness (kLuma for the SSE2-exact flavour, kLuma8 for the float-reference one);
// fbG / fbB / fbNa = the three lane biases of that flavour (BlockFront / RefFront).
template <bool kDxt, bool kEtc, uint32_t kWeights>
GB_DEV void selectors_from_flag_bytes(const uint32_t (&p)[16], uint32_t fbG, uint32_t fbB, uint32_t fbNa, uint32_t& dxtWord1,
                                      uint32_t& etcWord1)
{
    uint32_t idx = 0xFFu;
    uint32_t far0 = 0, far1 = 0, neg0 = 0xFFu, neg1 = 0xFFu;
    uint32_t gAbove[2] = {0, 0};
#pragma unroll
    for (int y = 3; y >= 0; --y) {
        uint32_t b[2], na[2], g[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // low lane = pixel x = 2h, high lane = pixel x = 2h + 1; lane = e + 0x8000, so bit 15 is G as it stands
            const uint32_t hi = dp4a(p[4 * y + 2 * h + 1], kWeights, fbG);
            g[h] = dp4a(p[4 * y + 2 * h], kWeights, hi * 65536u + fbG);
            b[h] = g[h] - fbB;
            na[h] = fbNa - g[h];
        }
        const uint32_t outerL = sign_bytes<0xFDB9>(b[0], na[0]);
        const uint32_t outerR = sign_bytes<0xFDB9>(b[1], na[1]);
        if (kEtc && !kDxt) {
            // ETC1s alone: a plane byte holds one PAIR column of all four rows, so the G flags are gathered per
            // pair from two rows at a time -- bytes (G(2h,y+1), G(2h+1,y+1), G(2h,y), G(2h+1,y)) -- one IDP each
            far1 = dp4a_su(outerL, kFlagWeights[3 + y], far1);
            far0 = dp4a_su(outerR, kFlagWeights[3 + y], far0);
            if (y & 1) {
                gAbove[0] = g[0];
                gAbove[1] = g[1];
            } else {
                neg1 = dp4a_su(sign_bytes<0xFDB9>(gAbove[0], g[0]), kFlagWeights[15 + (y >> 1)], neg1);
                neg0 = dp4a_su(sign_bytes<0xFDB9>(gAbove[1], g[1]), kFlagWeights[15 + (y >> 1)], neg0);
            }
            continue;
        }
        const uint32_t gez = sign_bytes<0xFDB9>(g[0], g[1]);
        if (kDxt) {
            if (y != 3) idx = idx * 256u + 0xFFu;
            idx = dp4a_su(gez, kFlagWeights[0], idx);
            idx = dp4a_su(outerL, kFlagWeights[1], idx);
            idx = dp4a_su(outerR, kFlagWeights[2], idx);
        }
        if (kEtc) {
            far1 = dp4a_su(outerL, kFlagWeights[3 + y], far1);
            far0 = dp4a_su(outerR, kFlagWeights[3 + y], far0);
            neg1 = dp4a_su(gez, kFlagWeights[7 + y], neg1);
            neg0 = dp4a_su(gez, kFlagWeights[11 + y], neg0);
        }
    }
    dxtWord1 = idx;
    // far0/far1 hold MINUS the plane bytes
    etcWord1 = (far1 * 256u + far0) * 0xFFFF0000u + (neg1 * 256u + neg0);
}

// ---------------------------------------------------------------------------------------- DXT1
// Output (goofy_tc.h:1276-1357): word0 = c0 | c1 << 16 with c0 = max corner 565 | 0x20,
// c1 = min corner 565 (both with the green LSB cleared); word1 = 2 bits per pixel, row major,
// bit0 = !Gez, bit1 = Lqt.
GB_DEV uint32_t dxt1_indices_lanes(const uint32_t (&p)[16], const BlockFront& f)
{
    // low lane walks pixels 0..7, high lane pixels 8..15; each step pushes 2 bits per lane
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t e = lanes_of(p[i], p[i + 8], f.laneBias);
        const uint32_t x = lqt_lanes(e, f);
        const uint32_t z = bitsel(x, ~e, 0x80008000u);  // bit15 = Lqt, bit14 = !Gez
        acc = bitsel(z, acc >> 2, 0xC000C000u);
    }
    return acc;
}

GB_DEV uint32_t dxt1_endpoints(const BlockFront& f)
{
    // 888 -> 565 for both corners at once: low lane = max corner, high lane = min corner.
    // to5 sits in the top 5 bits of a lane holding (max(v,1)-1) << 8.
    uint32_t rr = prmt(f.mxRB, f.mnRB, 0x5010);  // (0, maxR, 0, minR)
    uint32_t bb = prmt(f.mxRB, f.mnRB, 0x7030);  // (0, maxB, 0, minB)
    uint32_t gg = prmt(f.mxG, f.mnG, 0x5410);    // (., maxG, ., minG)
    rr = max2_u16x2(rr, 0x01000100u) - 0x01000100u;
    gg = max2_u16x2(gg, 0x01000100u) - 0x01000100u;
    // the odd constant also sets bit 16, which the >> 11 below turns into c0's forced green LSB (0x20)
    bb = max2_u16x2(bb, 0x01000100u) - 0x00FF0100u;
    const uint32_t rg = bitsel(rr, gg >> 5, 0xF800F800u);
    return bitsel(rg, bb >> 11, 0xFFC0FFC0u);
}

template <int SEL = kSelLanes>
GB_DEV void encode_dxt1(const uint32_t (&p)[16], const BlockFront& f, uint32_t& word0, uint32_t& word1)
{
    if (SEL == kSelFlagBytes) {
        uint32_t unused;
        selectors_from_flag_bytes<true, false, kLuma>(p, f.fbG, f.fbB, f.fbNa, word1, unused);
    } else {
        word1 = dxt1_indices_lanes(p, f);
    }
    word0 = dxt1_endpoints(f);
}

// ---------------------------------------------------------------------------------------- ETC1s
// floor((a+b)/2) on four bytes at once
GB_DEV uint32_t floor_avg4(uint32_t a, uint32_t b) { return (a & b) + (xor_and(a, b, 0xFEFEFEFEu) >> 1); }
// floor_avg4(~a, ~b) without materialising the complements
GB_DEV uint32_t floor_avg4_of_complements(uint32_t a, uint32_t b) { return nor(a, b) + (xor_and(a, b, 0xFEFEFEFEu) >> 1); }

// Output (goofy_tc.h:1358-1493): word0 = R5<<3 | G5<<11 | B5<<19 | control<<24;
// word1 = ~(GezPlane | LqtPlane << 16), pixel (x,y) at plane bit ((x^2)<<2)+y.
// `controlLut[range]` = control byte << 24 (the reference's table, goofy_tc.h:1040-1057).
GB_DEV uint32_t etc1_planes(const uint32_t (&p)[16], const BlockFront& f)
{
    uint32_t unused, planes;
    selectors_from_flag_bytes<false, true, kLuma>(p, f.fbG, f.fbB, f.fbNa, unused, planes);
    return planes;
}

// Average colour: the reference's fixed tree of rounded-UP averages (goofy_tc.h:1402-1414), evaluated on
// complemented bytes so each node is a floor average (3 ops).  First the part that reads the pixels: the four
// column averages (complemented) -- after it the pixels are dead.
GB_DEV void etc1_column_averages(const uint32_t (&p)[16], uint32_t (&col)[4])
{
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        const uint32_t top = floor_avg4_of_complements(p[x], p[4 + x]);
        const uint32_t bot = floor_avg4_of_complements(p[8 + x], p[12 + x]);
        col[x] = floor_avg4(top, bot);
    }
}

// word0 of an ETC1s block from the column averages: base colour 555 (in 888 positions) and the control byte
GB_DEV uint32_t etc1_base_word_from_columns(const uint32_t (&col)[4], uint32_t mid, uint32_t range, const uint32_t* controlLut)
{
    const uint32_t navg = floor_avg4(floor_avg4(col[0], col[1]), floor_avg4(col[2], col[3]));  // ~avg: bytes 255 - avg

    // Shift the average colour so its brightness becomes `mid` (goofy_tc.h:1431-1449):
    // base = clamp(avg + d, 0, 255) per channel with d = clamp(mid - Y(avg), -127, 127), then
    // to5(base) = (max(base,1) - 1) >> 3.  Both clamps fold into one: max(base,1) - 1 ==
    // clamp(avg + (d - 1), 0, 254), and "& 0xF8" leaves to5 << 3 in place.
    // The tree leaves the COMPLEMENT of the average and it is used as it stands: R+2G+B of the complement is
    // 1020 - Y4(avg), and (Y4 + 3) >> 2 == 255 - ((1020 - Y4) >> 2), so Y(avg) = 255 - (dp4a(navg) >> 2).
    int dm1 = (int)mid - 256 + (int)(dp4a(navg, kLuma, 0u) >> 2);   // d - 1 = mid - 1 - Y(avg)
    dm1 = dm1 < -128 ? -128 : dm1;
    dm1 = dm1 > 126 ? 126 : dm1;
    const uint32_t d2 = prmt((uint32_t)dm1, 0u, 0x1010);
    const uint32_t rb = addclamp_s16x2(~navg & 0x00FF00FFu, d2, 0x00FE00FEu);   // lanes (R, B)
    // green stays in bits 8..15: both addends are multiples of 256
    const uint32_t g = (uint32_t)addclamp_s32((int)(~navg & 0x0000FF00u), dm1 * 256, 254 * 256);
    return (rb & 0x00F800F8u) | (g & 0xF800u) | controlLut[range];
}

GB_DEV uint32_t etc1_base_word(const uint32_t (&p)[16], const BlockFront& f, const uint32_t* controlLut)
{
    uint32_t col[4];
    etc1_column_averages(p, col);
    return etc1_base_word_from_columns(col, f.mid, f.range, controlLut);
}

GB_DEV void encode_etc1(const uint32_t (&p)[16], const BlockFront& f, const uint32_t* controlLut, uint32_t& word0,
                        uint32_t& word1)
{
    word1 = etc1_planes(p, f);
    word0 = etc1_base_word(p, f, controlLut);
}

// Both codecs from one block: the twelve flag words are formed once and weighted twice.
GB_DEV void encode_both(const uint32_t (&p)[16], const BlockFront& f, const uint32_t* controlLut, uint32_t& dxt0,
                        uint32_t& dxt1, uint32_t& etc0, uint32_t& etc1)
{
    selectors_from_flag_bytes<true, true, kLuma>(p, f.fbG, f.fbB, f.fbNa, dxt1, etc1);
    dxt0 = dxt1_endpoints(f);
    etc0 = etc1_base_word(p, f, controlLut);
}

// ======================================================================== float-reference flavour
// goofyRef::compressDXT1/ETC1 (Src/goofy_tc_reference.cpp:514-623 goofyCompressBlock, :634-670 DXT1
// pack, :684-792 ETC1S pack) computes in float, but every value it forms is a multiple of 1/64 below
// 2^15, so it is reproduced here EXACTLY in integers (and is therefore immune to FMA contraction):
//   Y4 = R + 2G + B (= 4 * brightness)          range4 = max(maxY4 - minY4, 4 * minRange)
//   mid8 = maxY4 + minY4 (= 8 * midPoint)       q = 0.375 * range  ->  Q = 3 * range4 (= 32 q)
//   diff = Y - mid  ->  d = 8 * Y4 - 4 * mid8 (= 32 diff);   "diff > 0" <=> d >= 1;   "|diff| < q" <=> |d| < Q
// Results differ from the SSE2 flavour by design; this flavour exists so
// that "bit-exact against Src/goofy_tc_reference.cpp" can be met literally.
constexpr uint32_t kLuma8 = 0x00081008u;  // dp4a weights 8R + 16G + 8B = 8 * Y4

struct RefFront {
    uint32_t mnG, mxG, mnRB, mxRB;  // same layout as BlockFront
    uint32_t range4, mid8;
    // flag-byte scheme (selectors_from_flag_bytes): lane = d + 0x7FFF, so bit 15 <=> d >= 1 ("diff > 0");
    // bit 15 of (lane - fbB) <=> d >= Q; bit 15 of (fbNa - lane) <=> d <= -Q; near = neither of the last two
    uint32_t fbG, fbB, fbNa;
};

GB_DEV RefFront analyse_ref(const uint32_t (&p)[16], uint32_t minRange4)
{
    RefFront f;
    uint32_t q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) q[i] = p[i] << 8;
    f.mnG = reduce16<false>(p);
    f.mxG = reduce16<true>(p);
    f.mnRB = reduce16<false>(q);
    f.mxRB = reduce16<true>(q);
    const uint32_t minY4 = dp4a(prmt(f.mnRB, f.mnG, 0x3351), kLuma, 0u);
    const uint32_t maxY4 = dp4a(prmt(f.mxRB, f.mxG, 0x3351), kLuma, 0u);
    const uint32_t spread = maxY4 - minY4;
    f.range4 = spread > minRange4 ? spread : minRange4;   // :549
    f.mid8 = maxY4 + minY4;                                // :551
    const uint32_t Q = 3u * f.range4;                      // :554
    // |d| <= 8160 and 96 <= Q <= 3060: every biased lane stays inside 1..0xFFFE
    f.fbG = 0x7FFFu - 4u * f.mid8;
    f.fbB = (Q - 1u) * 0x10001u;
    f.fbNa = (0xFFFFu - Q) * 0x10001u;
    return f;
}

// :634-670  c0 = (R>>3)<<11 | (G>>3)<<6 | B>>3 | 0x20 from the max corner, c1 from the min corner
GB_DEV void encode_dxt1_ref(const uint32_t (&p)[16], const RefFront& f, uint32_t& word0, uint32_t& word1)
{
    uint32_t unused;
    selectors_from_flag_bytes<true, false, kLuma8>(p, f.fbG, f.fbB, f.fbNa, word1, unused);   // bit1 = near, bit0 = !(diff > 0)
    const uint32_t rr = prmt(f.mxRB, f.mnRB, 0x5010);
    const uint32_t bb = prmt(f.mxRB, f.mnRB, 0x7030);
    const uint32_t gg = prmt(f.mxG, f.mnG, 0x5410);
    const uint32_t rg = bitsel(rr, gg >> 5, 0xF800F800u);
    word0 = (bitsel(rg, bb >> 11, 0xFFC0FFC0u) & 0xFFDFFFDFu) | 0x20u;
}

// :758-792  base colour = average colour moved onto the mid brightness, (v & 0xF8) packing,
// control byte from floatToByte(range / 2) against {10,21,36,52,75,90,126} (`controlLut[brightRange]`).
GB_DEV void encode_etc1_ref(const uint32_t (&p)[16], const RefFront& f, const uint32_t* controlLut, uint32_t& word0,
                            uint32_t& word1)
{
    uint32_t unused;
    selectors_from_flag_bytes<false, true, kLuma8>(p, f.fbG, f.fbB, f.fbNa, unused, word1);   // !(diff > 0) plane | far plane << 16

    // :526-536 (avgColor * 16).  Two channel sums per instruction: IDP.2A with the 16-bit weights (1, 4096) adds
    // R + 4096 G (resp. B + 4096 A) of a pixel; sixteen bytes sum to at most 4080 < 4096, so the fields never meet.
    uint32_t sumRG = 0, sumBA = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        sumRG = dp2a_lo(0x10000001u, p[i], sumRG);
        sumBA = dp2a_hi(0x10000001u, p[i], sumBA);
    }
    const uint32_t sumR = sumRG & 0xFFFu, sumG = sumRG >> 12, sumB = sumBA & 0xFFFu;
    // everything * 64: avg = sum * 4, avgY = sumR + 2 sumG + sumB, mid = mid8 * 8; floatToByte = floor(v + 0.5) clamped
    const int diffY64 = (int)(8u * f.mid8) - (int)(sumR + 2u * sumG + sumB);   // :606-607
    const int sums[3] = {(int)sumR, (int)sumG, (int)sumB};
    uint32_t packed = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int v = (4 * sums[c] + diffY64 + 32) >> 6;   // arithmetic shift = floor; negative -> clamped to 0 below
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        packed |= ((uint32_t)v & 0xF8u) << (8 * c);
    }
    const uint32_t brightRange = (f.range4 + 4u) >> 3;   // :603 floatToByte(range * 0.5)
    word0 = packed | controlLut[brightRange];
}

// :684-720 getEtc1SBlockControlByte
GB_DEV uint32_t etc1_control_word_ref(uint32_t brightRange)
{
    const uint32_t cw = (brightRange > 10u) + (brightRange > 21u) + (brightRange > 36u) + (brightRange > 52u) +
                        (brightRange > 75u) + (brightRange > 90u) + (brightRange > 126u);
    return (cw * 36u + 3u) << 24;
}

// The reference's table (goofy_tc.h:1040-1057) steps at 22,44,74,106,152,182,254 and holds
// (cw<<5 | cw<<2 | 0b11) << 24: both sub-block tables equal, diff = 1, flip = 1.
GB_DEV uint32_t etc1_control_word(uint32_t range)
{
    const uint32_t cw = (range >= 22u) + (range >= 44u) + (range >= 74u) + (range >= 106u) + (range >= 152u) +
                        (range >= 182u) + (range >= 254u);
    return (cw * 36u + 3u) << 24;
}

}  // namespace gb
