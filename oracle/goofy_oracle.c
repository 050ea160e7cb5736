/*
 * goofy_oracle.c -- scalar CPU oracle (TEST INFRASTRUCTURE ONLY, see goofy_oracle.h).
 *
 * Written from scratch as a per-block, per-lane restatement of the reference's
 * SSE2 tile kernel.  The reference works on 16 byte lanes at a time over a
 * 16x4-pixel tile; here every lane operation is applied to one scalar byte of
 * one 4x4 block, keeping the same order of rounded operations so the rounding
 * cascades are preserved.  No closed forms are used on purpose: the CUDA kernel
 * uses closed forms, and the tests prove the two agree.
 */
#include "goofy_oracle.h"

#include <string.h>

/* ---- byte-lane primitives (reference: GoofyTC/goofy_tc.h:645-654, 719-744, 666-690) ---- */
static inline unsigned u8_avg(unsigned a, unsigned b) { return (a + b + 1u) >> 1; }       /* pavgb   */
static inline unsigned u8_subsat(unsigned a, unsigned b) { return a > b ? a - b : 0u; }     /* psubusb */
static inline unsigned u8_addsat(unsigned a, unsigned b) { unsigned s = a + b; return s > 255u ? 255u : s; } /* paddusb */
static inline unsigned u8_min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned u8_max(unsigned a, unsigned b) { return a > b ? a : b; }

/* Perceptual brightness as the reference computes it: avg(avg(R,B),G)  (goofy_tc.h:1166,1205) */
static inline unsigned luma(unsigned r, unsigned g, unsigned b) { return u8_avg(u8_avg(r, b), g); }

/* 8-bit -> 5-bit as the reference does it: three rounded halvings after subtracting 8
 * (goofy_tc.h:1309-1311, 1462) */
static inline unsigned to5(unsigned v)
{
    return u8_avg(u8_avg(u8_avg(u8_subsat(v, 8u), 0u), 0u), 0u);
}

/* ETC1 control byte from the clamped brightness range (table at goofy_tc.h:1040-1057).
 * The table holds 0xTT000000 with TT = cw<<5 | cw<<2 | diff(1)<<1 | flip(1); its steps
 * are at 22, 44, 74, 106, 152, 182, 254 (read off the table rows). */
static inline unsigned etc1_control_byte(unsigned range)
{
    static const unsigned step[7] = {22u, 44u, 74u, 106u, 152u, 182u, 254u};
    unsigned cw = 0;
    while (cw < 7u && range >= step[cw]) ++cw;
    return (cw << 5) | (cw << 2) | 3u;
}

typedef struct {
    unsigned mn[3], mx[3];   /* bounding box corners (goofy_tc.h:1103-1140) */
    unsigned range, mid, qt; /* goofy_tc.h:1174-1190 */
    unsigned gez[4][4];      /* [y][x] 1 if pixel brightness >= mid  (goofy_tc.h:1215) */
    unsigned lqt[4][4];      /* [y][x] 1 if |brightness - mid| < qt   (goofy_tc.h:1224) */
} block_state;

static void analyse_block(const uint8_t* px, size_t stride, block_state* s)
{
    for (int c = 0; c < 3; ++c) { s->mn[c] = 255u; s->mx[c] = 0u; }
    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x)
            for (int c = 0; c < 3; ++c) {
                unsigned v = px[(size_t)y * stride + 4u * (unsigned)x + (unsigned)c];
                s->mn[c] = u8_min(s->mn[c], v);
                s->mx[c] = u8_max(s->mx[c], v);
            }

    /* brightness of the two box corners (not of real pixels) */
    unsigned minY = luma(s->mn[0], s->mn[1], s->mn[2]);
    unsigned maxY = luma(s->mx[0], s->mx[1], s->mx[2]);

    s->range = u8_max(u8_subsat(maxY, minY), 8u);
    s->mid = u8_avg(minY, maxY);
    unsigned half = u8_avg(s->range, 0u);
    unsigned quarter = u8_avg(half, 0u);
    unsigned eighth = u8_avg(quarter, 0u);
    s->qt = u8_addsat(quarter, eighth);

    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) {
            const uint8_t* p = px + (size_t)y * stride + 4u * (unsigned)x;
            unsigned Y = luma(p[0], p[1], p[2]);
            unsigned pos = u8_min(u8_subsat(Y, s->mid), 127u);
            unsigned neg = u8_min(u8_subsat(s->mid, Y), 127u);
            unsigned absd = pos | neg;
            s->gez[y][x] = (neg == 0u);
            /* pcmpgtb is a signed compare; both operands are <= 127 here so it is a plain '<' */
            s->lqt[y][x] = ((int8_t)absd < (int8_t)s->qt);
        }
}

static inline void store_le32(uint8_t* p, uint32_t v)
{
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

void goofy_oracle_block_dxt1(const uint8_t* px, size_t stride, uint8_t* out)
{
    block_state s;
    analyse_block(px, stride, &s);

    /* goofy_tc.h:1292-1301: per pixel, bit0 = NOT gez, bit1 = lqt; row-major, 2 bits each */
    uint32_t indices = 0;
    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) {
            uint32_t code = (s.gez[y][x] ? 0u : 1u) | (s.lqt[y][x] ? 2u : 0u);
            indices |= code << (2 * (4 * y + x));
        }

    /* goofy_tc.h:1334-1336: colour0 = max corner as R5:G5:0:B5 with the green LSB forced to 1
     * (keeps colour0 > colour1, i.e. 4-colour mode), colour1 = min corner */
    uint32_t c0 = (to5(s.mx[0]) << 11) | (to5(s.mx[1]) << 6) | to5(s.mx[2]) | 0x20u;
    uint32_t c1 = (to5(s.mn[0]) << 11) | (to5(s.mn[1]) << 6) | to5(s.mn[2]);
    store_le32(out, c0 | (c1 << 16));
    store_le32(out + 4, indices);
}

void goofy_oracle_block_etc1(const uint8_t* px, size_t stride, uint8_t* out)
{
    block_state s;
    analyse_block(px, stride, &s);

    /* Selector planes.  The reference transposes the masks to column-major with
     * the column pairs swapped (goofy_tc.h:1366-1375) because it stores a
     * big-endian 16-bit field through a little-endian dword: pixel (x,y) lands at
     * bit ((x^2)<<2)+y of each 16-bit plane. */
    uint32_t pos_or_zero = 0, less_than_qt = 0;
    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) {
            unsigned bit = (unsigned)(((x ^ 2) << 2) + y);
            pos_or_zero |= (uint32_t)s.gez[y][x] << bit;
            less_than_qt |= (uint32_t)s.lqt[y][x] << bit;
        }
    uint32_t word1 = ~(pos_or_zero | (less_than_qt << 16));   /* goofy_tc.h:1479 */

    /* Average colour: a fixed tree of rounded-up averages, rows (0,1) and (2,3)
     * per column first, then columns (0,1) and (2,3)  (goofy_tc.h:1402-1414). */
    unsigned avg[3];
    for (int c = 0; c < 3; ++c) {
        unsigned col[4];
        for (int x = 0; x < 4; ++x) {
            unsigned r0 = px[0 * stride + 4u * (unsigned)x + (unsigned)c];
            unsigned r1 = px[1 * stride + 4u * (unsigned)x + (unsigned)c];
            unsigned r2 = px[2 * stride + 4u * (unsigned)x + (unsigned)c];
            unsigned r3 = px[3 * stride + 4u * (unsigned)x + (unsigned)c];
            col[x] = u8_avg(u8_avg(r0, r1), u8_avg(r2, r3));
        }
        avg[c] = u8_avg(u8_avg(col[0], col[1]), u8_avg(col[2], col[3]));
    }

    /* Move the average colour's brightness onto the block mid brightness while
     * keeping its chroma (goofy_tc.h:1431-1449). */
    unsigned avgY = luma(avg[0], avg[1], avg[2]);
    unsigned pos_corr = u8_min(u8_subsat(s.mid, avgY), 127u);
    unsigned neg_corr = u8_min(u8_subsat(avgY, s.mid), 127u);
    unsigned up = (neg_corr == 0u);
    unsigned corr = pos_corr | neg_corr;
    unsigned base[3];
    for (int c = 0; c < 3; ++c)
        base[c] = up ? u8_addsat(avg[c], corr) : u8_subsat(avg[c], corr);

    /* goofy_tc.h:1462-1478: byte0 = R5<<3, byte1 = G5<<3, byte2 = B5<<3 (delta bits 0), byte3 = control */
    uint32_t word0 = (to5(base[0]) << 3) | (to5(base[1]) << 11) | (to5(base[2]) << 19) |
                     (etc1_control_byte(s.range) << 24);
    store_le32(out, word0);
    store_le32(out + 4, word1);
}

static int compress_image(uint8_t* result, const uint8_t* input, unsigned width, unsigned height,
                          unsigned stride, void (*block_fn)(const uint8_t*, size_t, uint8_t*))
{
    /* goofy_tc.h:1500-1508 */
    if (width % 16u != 0u) return -1;
    if (height % 4u != 0u) return -2;
    unsigned bw = width >> 2, bh = height >> 2;
    for (unsigned by = 0; by < bh; ++by)
        for (unsigned bx = 0; bx < bw; ++bx) {
            block_fn(input + (size_t)by * 4u * stride + (size_t)bx * 16u, stride, result);
            result += 8;
        }
    return 0;
}

int goofy_oracle_compress_dxt1(uint8_t* result, const uint8_t* input, unsigned width, unsigned height, unsigned stride)
{
    return compress_image(result, input, width, height, stride, goofy_oracle_block_dxt1);
}

int goofy_oracle_compress_etc1(uint8_t* result, const uint8_t* input, unsigned width, unsigned height, unsigned stride)
{
    return compress_image(result, input, width, height, stride, goofy_oracle_block_etc1);
}

/* ------------------------------------------------------------------ float-reference flavour
 * Restates goofyCompressBlock and its two packers with the same float expressions
 * (Src/goofy_tc_reference.cpp:514-623).  Every intermediate is a multiple of 1/64 below 2^15,
 * so float evaluation is exact and compiler contraction cannot change the result. */
typedef struct {
    unsigned mn[3], mx[3], base[3];
    unsigned bright_range;
    int index[16];
} floatref_block;

static unsigned float_to_byte(float v) /* :350-361 */
{
    v = v + 0.5f;
    if (v < 0.0f) v = 0.0f;
    if (v > 255.0f) v = 255.0f;
    return (unsigned)(uint8_t)v;
}

static float brightness_f(float r, float g, float b) { return r * 0.25f + g * 0.5f + b * 0.25f; } /* :246-250 */

static void floatref_analyse(const uint8_t* px, size_t stride, float min_range, floatref_block* o)
{
    float mn[3] = {99999.0f, 99999.0f, 99999.0f}, mx[3] = {-99999.0f, -99999.0f, -99999.0f}, avg[3] = {0, 0, 0};
    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x)
            for (int c = 0; c < 3; ++c) {
                float v = (float)px[(size_t)y * stride + 4u * (unsigned)x + (unsigned)c];
                if (v < mn[c]) mn[c] = v;
                if (v > mx[c]) mx[c] = v;
                avg[c] += v;
            }
    for (int c = 0; c < 3; ++c) avg[c] /= 16.0f;
    float maxY = brightness_f(mx[0], mx[1], mx[2]), minY = brightness_f(mn[0], mn[1], mn[2]);
    float range = maxY - minY;
    if (range < min_range) range = min_range;                       /* :549 */
    float mid = (maxY + minY) * 0.5f;                               /* :551 */
    float q = range * 0.375f;                                       /* :554 */
    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) {
            const uint8_t* p = px + (size_t)y * stride + 4u * (unsigned)x;
            float diff = brightness_f((float)p[0], (float)p[1], (float)p[2]) - mid;
            float ad = diff < 0.0f ? -diff : diff;
            o->index[4 * y + x] = diff > 0.0f ? (diff < q ? 2 : 0) : (ad < q ? 3 : 1);   /* :579-589 */
        }
    for (int c = 0; c < 3; ++c) {
        o->mn[c] = float_to_byte(mn[c]);
        o->mx[c] = float_to_byte(mx[c]);
    }
    o->bright_range = float_to_byte(range * 0.5f);                  /* :603 */
    float avgY = brightness_f(avg[0], avg[1], avg[2]);
    float diffY = mid - avgY;                                       /* :606-608 */
    for (int c = 0; c < 3; ++c) o->base[c] = float_to_byte(avg[c] + diffY);
}

static void floatref_block_dxt1(const uint8_t* px, size_t stride, uint8_t* out)
{
    floatref_block b;
    floatref_analyse(px, stride, 8.0f, &b);                         /* :667 */
    uint32_t c0 = ((b.mx[0] >> 3) << 11) | ((b.mx[1] >> 3) << 6) | (b.mx[2] >> 3) | 0x20u;   /* :634-646, :670 */
    uint32_t c1 = ((b.mn[0] >> 3) << 11) | ((b.mn[1] >> 3) << 6) | (b.mn[2] >> 3);
    uint32_t idx = 0;
    for (int n = 0; n < 16; ++n) idx |= (uint32_t)(b.index[n] & 3) << (2 * n);              /* :648-657 */
    store_le32(out, c0 | (c1 << 16));
    store_le32(out + 4, idx);
}

static void floatref_block_etc1(const uint8_t* px, size_t stride, uint8_t* out)
{
    static const unsigned table[7] = {10u, 21u, 36u, 52u, 75u, 90u, 126u};                  /* :688 */
    floatref_block b;
    floatref_analyse(px, stride, 16.0f, &b);                        /* :766 */
    unsigned cw = 0;
    while (cw < 7u && b.bright_range > table[cw]) ++cw;
    uint32_t word0 = (b.base[0] & 0xF8u) | ((b.base[1] & 0xF8u) << 8) | ((b.base[2] & 0xF8u) << 16) |   /* :625-632 */
                     (((cw << 5) | (cw << 2) | 3u) << 24);
    uint32_t v = 0;
    for (int n = 0; n < 16; ++n) {                                  /* :722-756 */
        int x = n & 3, y = n >> 2;
        unsigned bit = (unsigned)(((x ^ 2) << 2) + y);
        unsigned far_ = (b.index[n] == 0 || b.index[n] == 1), neg = (b.index[n] == 1 || b.index[n] == 3);
        v |= (uint32_t)neg << bit;
        v |= (uint32_t)far_ << (bit + 16);
    }
    store_le32(out, word0);
    store_le32(out + 4, v);
}

static int floatref_compress(uint8_t* result, const uint8_t* input, unsigned width, unsigned height, unsigned stride,
                             void (*block_fn)(const uint8_t*, size_t, uint8_t*))
{
    if (width % 4u != 0u) return -1;                                /* :796-804 */
    if (height % 4u != 0u) return -2;
    for (unsigned by = 0; by < height / 4u; ++by)
        for (unsigned bx = 0; bx < width / 4u; ++bx) {
            block_fn(input + (size_t)by * 4u * stride + (size_t)bx * 16u, stride, result);
            result += 8;
        }
    return 0;
}

int goofy_oracle_floatref_compress_dxt1(uint8_t* result, const uint8_t* input, unsigned width, unsigned height, unsigned stride)
{
    return floatref_compress(result, input, width, height, stride, floatref_block_dxt1);
}

int goofy_oracle_floatref_compress_etc1(uint8_t* result, const uint8_t* input, unsigned width, unsigned height, unsigned stride)
{
    return floatref_compress(result, input, width, height, stride, floatref_block_etc1);
}

/* ------------------------------------------------------------------ decoders */

static inline uint32_t load_le32(const uint8_t* p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* BC1 colour block (Src/decoder.cpp:798-871): RGB565 endpoints expanded by bit
 * replication; c0 > c1 selects the 4-colour mode with truncating thirds. */
static void decode_block_dxt1(const uint8_t* blk, uint8_t* dst, size_t dst_stride)
{
    unsigned v0 = blk[0] | (blk[1] << 8), v1 = blk[2] | (blk[3] << 8);
    unsigned pal[4][4];
    unsigned e[2] = {v0, v1};
    for (int k = 0; k < 2; ++k) {
        unsigned r = (e[k] >> 11) & 31u, g = (e[k] >> 5) & 63u, b = e[k] & 31u;
        pal[k][0] = (r << 3) | (r >> 2);
        pal[k][1] = (g << 2) | (g >> 4);
        pal[k][2] = (b << 3) | (b >> 2);
        pal[k][3] = 255u;
    }
    for (int c = 0; c < 3; ++c) {
        if (v0 <= v1) {
            pal[2][c] = (pal[0][c] + pal[1][c]) / 2u;
            pal[3][c] = 0u;
        } else {
            pal[2][c] = (2u * pal[0][c] + pal[1][c]) / 3u;
            pal[3][c] = (pal[0][c] + 2u * pal[1][c]) / 3u;
        }
    }
    pal[2][3] = 255u;
    pal[3][3] = (v0 <= v1) ? 0u : 255u;
    uint32_t idx = load_le32(blk + 4);
    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) {
            unsigned k = (idx >> (2 * (4 * y + x))) & 3u;
            for (int c = 0; c < 4; ++c) dst[(size_t)y * dst_stride + 4u * (unsigned)x + (unsigned)c] = (uint8_t)pal[k][c];
        }
}

static inline unsigned clamp255(int v) { return v < 0 ? 0u : (v > 255 ? 255u : (unsigned)v); }

/* ETC1 block (Src/decoder.cpp:388-678 for the differential mode the encoder emits;
 * the individual mode is included so the decoder is a complete ETC1 decoder). */
static void decode_block_etc1(const uint8_t* blk, uint8_t* dst, size_t dst_stride)
{
    static const int modifier[8][2] = {{2, 8}, {5, 17}, {9, 29}, {13, 42}, {18, 60}, {24, 80}, {33, 106}, {47, 183}};
    unsigned diff = (blk[3] >> 1) & 1u, flip = blk[3] & 1u;
    unsigned cw[2] = {(unsigned)(blk[3] >> 5) & 7u, (unsigned)(blk[3] >> 2) & 7u};
    unsigned base[2][3];
    for (int c = 0; c < 3; ++c) {
        if (diff) {
            unsigned c5 = blk[c] >> 3;
            int d = (int)(blk[c] & 7u);
            if (d >= 4) d -= 8;
            unsigned c5b = (unsigned)((int)c5 + d) & 31u;
            base[0][c] = (c5 << 3) | (c5 >> 2);
            base[1][c] = (c5b << 3) | (c5b >> 2);
        } else {
            unsigned a = blk[c] >> 4, b = blk[c] & 15u;
            base[0][c] = a * 17u;
            base[1][c] = b * 17u;
        }
    }
    unsigned msb = ((unsigned)blk[4] << 8) | blk[5];
    unsigned lsb = ((unsigned)blk[6] << 8) | blk[7];
    for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 4; ++y) {
            unsigned bit = (unsigned)(4 * x + y);
            unsigned m = (msb >> bit) & 1u, l = (lsb >> bit) & 1u;
            int sub = flip ? (y >= 2) : (x >= 2);
            int mag = modifier[cw[sub]][l];
            int delta = m ? -mag : mag;
            uint8_t* o = dst + (size_t)y * dst_stride + 4u * (unsigned)x;
            for (int c = 0; c < 3; ++c) o[c] = (uint8_t)clamp255((int)base[sub][c] + delta);
            o[3] = 255u;
        }
}

static void decode_image(const uint8_t* blocks, unsigned width, unsigned height, uint8_t* rgba,
                         void (*fn)(const uint8_t*, uint8_t*, size_t))
{
    unsigned bw = width >> 2, bh = height >> 2;
    size_t stride = (size_t)width * 4u;
    for (unsigned by = 0; by < bh; ++by)
        for (unsigned bx = 0; bx < bw; ++bx) {
            fn(blocks, rgba + (size_t)by * 4u * stride + (size_t)bx * 16u, stride);
            blocks += 8;
        }
}

void goofy_oracle_decode_dxt1(const uint8_t* blocks, unsigned width, unsigned height, uint8_t* rgba)
{
    decode_image(blocks, width, height, rgba, decode_block_dxt1);
}

void goofy_oracle_decode_etc1(const uint8_t* blocks, unsigned width, unsigned height, uint8_t* rgba)
{
    decode_image(blocks, width, height, rgba, decode_block_etc1);
}

void goofy_oracle_sse_rgb(const uint8_t* a, const uint8_t* b, size_t pixels, uint64_t sse[3])
{
    sse[0] = sse[1] = sse[2] = 0;
    for (size_t i = 0; i < pixels; ++i)
        for (int c = 0; c < 3; ++c) {
            int d = (int)a[4 * i + (size_t)c] - (int)b[4 * i + (size_t)c];
            sse[c] += (uint64_t)(d * d);
        }
}
