/*
 * goofy_oracle.h -- CPU oracle for the Goofy DXT1 / ETC1s block encoders.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain scalar C restatement of what the
 * reference's SSE2 path computes, one 4x4 block at a time.  It exists so the
 * CUDA path can be checked bit-for-bit.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * library (libgoofy_b200.so) never links or calls anything in oracle/.
 *
 * Parity pin: the oracle is checked against (a) the unmodified reference
 * compiled from /root/reference into oracle/_ref/libgoofy_ref.so (see
 * oracle/Makefile) and (b) the committed golden vectors in tests/golden/
 * that were produced by that library (tests/golden/make_golden.py).
 *
 * Reference semantics followed (all citations into /root/reference):
 *   GoofyTC/goofy_tc.h:1069-1494  goofySimdEncode<DXT1|ETC1>
 *   GoofyTC/goofy_tc.h:1497-1557  compressDXT1 / compressETC1 (shape checks, loops)
 *   GoofyTC/goofy_tc.h:645-654    avg  = (a+b+1)>>1
 *   GoofyTC/goofy_tc.h:719-744    addsatu / subsatu
 *   GoofyTC/goofy_tc.h:1040-1057  ETC1 range -> control byte table
 *   Src/decoder.cpp:798-871       BC1 colour decode
 *   Src/decoder.cpp:388-678       ETC1 differential decode
 *   Src/main.cpp:403-469          MSE / PSNR definition
 */
#ifndef GOOFY_ORACLE_H
#define GOOFY_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Encode one 4x4 block. `px` points at the top-left pixel (RGBA8), rows are
 * `stride` bytes apart.  Writes 8 bytes to `out`. */
void goofy_oracle_block_dxt1(const uint8_t* px, size_t stride, uint8_t* out);
void goofy_oracle_block_etc1(const uint8_t* px, size_t stride, uint8_t* out);

/* Whole image; same contract and return codes as goofy::compressDXT1/ETC1
 * (goofy_tc.h:1497-1557): 0 ok, -1 width%16, -2 height%4. */
int goofy_oracle_compress_dxt1(uint8_t* result, const uint8_t* input,
                               unsigned width, unsigned height, unsigned stride);
int goofy_oracle_compress_etc1(uint8_t* result, const uint8_t* input,
                               unsigned width, unsigned height, unsigned stride);

/* Second flavour: the reference's float "idea" encoder goofyRef::compressDXT1/ETC1
 * (Src/goofy_tc_reference.cpp:514-623, :634-670, :684-792).  It is NOT bit-equal to the SSE2
 * path (different rounding, tie-break, minimum range and table thresholds); it is restated
 * here with the same float expressions.  Shapes: width%4 (-1), height%4 (-2) like :794-850.
 * Block rows advance by `stride` (the reference advances by width*16 bytes, which is the same
 * thing for the tight stride its harness always uses). */
int goofy_oracle_floatref_compress_dxt1(uint8_t* result, const uint8_t* input,
                                        unsigned width, unsigned height, unsigned stride);
int goofy_oracle_floatref_compress_etc1(uint8_t* result, const uint8_t* input,
                                        unsigned width, unsigned height, unsigned stride);

/* Decode whole images of 8-byte blocks (row-major block order) to tight RGBA8.
 * Alpha is written as 255 (BC1 3-colour "transparent" index writes 0,0,0,0
 * like Src/decoder.cpp:836-851). */
void goofy_oracle_decode_dxt1(const uint8_t* blocks, unsigned width, unsigned height, uint8_t* rgba);
void goofy_oracle_decode_etc1(const uint8_t* blocks, unsigned width, unsigned height, uint8_t* rgba);

/* Sum of squared errors per channel (R,G,B) between two tight RGBA8 images.
 * PSNR is derived by the caller: psnrRGB768 = 10*log10(768^2 / ((sseR+sseG+sseB)/n))
 * (Src/main.cpp:444,466), textbook = 10*log10(255^2 / ((sseR+sseG+sseB)/(3n))). */
void goofy_oracle_sse_rgb(const uint8_t* a, const uint8_t* b, size_t pixels, uint64_t sse[3]);

#ifdef __cplusplus
}
#endif
#endif
