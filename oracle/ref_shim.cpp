// ref_shim.cpp -- C-ABI wrapper around the UNMODIFIED reference, for tests and the CPU baseline.
//
// TEST / BASELINE INFRASTRUCTURE ONLY.  This file contains no encoder logic of its
// own: it includes the reference header from where it lies (/root/reference, via -I)
// and is linked with the reference's float encoder and decoders, compiled in place by
// oracle/Makefile into oracle/_ref/libgoofy_ref.so.  No reference source is copied into
// this repository; oracle/_ref/ is git-ignored and travels to the GPU box as a binary.
//
// What it exposes:
//   ref_goofy_*       goofy::compressDXT1 / compressETC1     (GoofyTC/goofy_tc.h:1497-1557, SSE2 path)
//   ref_goofyref_*    goofyRef::compressDXT1 / compressETC1  (Src/goofy_tc_reference.cpp:794-850)
//   ref_decode_*      DecoderBC::decodeBlockDXT1 / ETC1      (Src/decoder.cpp:933-971) over a whole image
//   ref_goofy_compress_mt   row-parallel driver: the same goofy:: function called on T
//                           contiguous strips of block rows from T threads (the function is
//                           pure, goofy_tc.h has no mutable globals) -- the "all host cores"
//                           CPU baseline BASELINE.md section 4 describes.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "goofy_tc.h"            // -I/root/reference/GoofyTC  (defines goofy::compress*, one TU only)
#include "goofy_tc_reference.h"  // -I/root/reference/Src
#include "decoder.h"             // -I/root/reference/Src

extern "C" {

int ref_goofy_compress_dxt1(unsigned char* result, const unsigned char* input, unsigned width, unsigned height, unsigned stride)
{
    return goofy::compressDXT1(result, input, width, height, stride);
}

int ref_goofy_compress_etc1(unsigned char* result, const unsigned char* input, unsigned width, unsigned height, unsigned stride)
{
    return goofy::compressETC1(result, input, width, height, stride);
}

int ref_goofyref_compress_dxt1(unsigned char* result, const unsigned char* input, unsigned width, unsigned height, unsigned stride)
{
    return goofyRef::compressDXT1(result, input, width, height, stride);
}

int ref_goofyref_compress_etc1(unsigned char* result, const unsigned char* input, unsigned width, unsigned height, unsigned stride)
{
    return goofyRef::compressETC1(result, input, width, height, stride);
}

void ref_decode_dxt1(const unsigned char* blocks, unsigned width, unsigned height, unsigned char* rgba)
{
    const size_t stride = (size_t)width * 4;
    for (unsigned by = 0; by < height / 4; ++by)
        for (unsigned bx = 0; bx < width / 4; ++bx, blocks += 8)
            DecoderBC::decodeBlockDXT1(blocks, rgba + (size_t)by * 4 * stride + (size_t)bx * 16, stride);
}

void ref_decode_etc1(const unsigned char* blocks, unsigned width, unsigned height, unsigned char* rgba)
{
    const size_t stride = (size_t)width * 4;
    for (unsigned by = 0; by < height / 4; ++by)
        for (unsigned bx = 0; bx < width / 4; ++bx, blocks += 8)
            DecoderBC::decodeBlockETC1(blocks, rgba + (size_t)by * 4 * stride + (size_t)bx * 16, stride);
}

// codec: 0 = DXT1, 1 = ETC1.  threads <= 1 calls the function once, exactly as shipped.
int ref_goofy_compress_mt(int codec, unsigned char* result, const unsigned char* input, unsigned width,
                          unsigned height, unsigned stride, int threads)
{
    typedef int (*fn_t)(unsigned char*, const unsigned char*, unsigned, unsigned, unsigned);
    fn_t fn = codec == 0 ? goofy::compressDXT1 : goofy::compressETC1;
    if (threads <= 1 || (width % 16) != 0 || (height % 4) != 0 || height == 0) return fn(result, input, width, height, stride);
    const unsigned block_rows = height / 4;
    const unsigned T = (unsigned)threads < block_rows ? (unsigned)threads : block_rows;
    std::vector<std::thread> pool;
    std::vector<int> rc(T, 0);
    for (unsigned t = 0; t < T; ++t) {
        const unsigned r0 = (unsigned)((uint64_t)block_rows * t / T), r1 = (unsigned)((uint64_t)block_rows * (t + 1) / T);
        pool.emplace_back([=, &rc]() {
            rc[t] = fn(result + (size_t)r0 * (width / 4) * 8, input + (size_t)r0 * 4 * stride, width, (r1 - r0) * 4, stride);
        });
    }
    for (auto& th : pool) th.join();
    for (unsigned t = 0; t < T; ++t)
        if (rc[t]) return rc[t];
    return 0;
}

unsigned ref_hardware_threads(void) { return std::thread::hardware_concurrency(); }

}  // extern "C"
