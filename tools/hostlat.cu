// hostlat.cu -- where the time of a SMALL drop-in host call goes (768 x 512 by default; the reference harness times
// exactly this call per test image, Src/main.cpp:653-664 of the reference).  Diagnostic, not part of the product.
//
//   hostlat [width height] [iters]
//
// Prints best / median microseconds of:
//   memcpy        the staging copy alone: pageable -> pinned, one thread (what a pageable caller cannot avoid)
//   bare          pinned H2D -> encode kernel -> D2H -> cudaStreamSynchronize, issued by this program (the floor the
//                 library's pinned path can reach)
//   h2d / d2h     the two copies alone (pinned), each followed by a synchronize
//   lib pinned    goofy_b200_compress_dxt1 on pinned buffers
//   lib pageable  goofy_b200_compress_dxt1 on malloc'ed buffers (64-byte aligned)
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/goofy_b200.h"

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            std::fprintf(stderr, "%s:%d CUDA error %s\n", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            std::exit(2);                                                                               \
        }                                                                                               \
    } while (0)

template <typename F>
static void timeit(const char* name, uint32_t iters, double px, F fn)
{
    for (int i = 0; i < 5; ++i) fn();
    std::vector<double> us;
    for (uint32_t i = 0; i < iters; ++i) {
        const auto a = std::chrono::steady_clock::now();
        fn();
        us.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count());
    }
    std::sort(us.begin(), us.end());
    std::printf("%-14s best %8.1f us   median %8.1f us   (%7.0f MP/s at best)\n", name, us.front(), us[us.size() / 2], px / us.front());
}

int main(int argc, char** argv)
{
    const uint32_t W = argc > 2 ? (uint32_t)std::atoi(argv[1]) : 768u, H = argc > 2 ? (uint32_t)std::atoi(argv[2]) : 512u;
    const uint32_t iters = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 300u;
    const size_t inBytes = (size_t)W * H * 4, outBytes = (size_t)W * H / 2;
    CK(cudaSetDevice(0));
    uint8_t *pinIn, *pinOut, *dIn, *dOut;
    CK(cudaHostAlloc(&pinIn, inBytes, cudaHostAllocDefault));
    CK(cudaHostAlloc(&pinOut, outBytes, cudaHostAllocDefault));
    CK(cudaMalloc(&dIn, inBytes));
    CK(cudaMalloc(&dOut, outBytes));
    uint8_t* pgIn = (uint8_t*)std::aligned_alloc(64, inBytes);
    uint8_t* pgOut = (uint8_t*)std::aligned_alloc(64, outBytes);
    uint8_t* want = (uint8_t*)std::malloc(outBytes);
    for (size_t i = 0; i < inBytes; ++i) pgIn[i] = (uint8_t)((i * 2654435761u) >> 13);
    std::memcpy(pinIn, pgIn, inBytes);
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    const double px = (double)W * H;
    std::printf("%u x %u RGBA8: %.0f KiB in, %.0f KiB out\n", W, H, inBytes / 1024.0, outBytes / 1024.0);

    timeit("memcpy", iters, px, [&] { std::memcpy(pinIn, pgIn, inBytes); });
    timeit("h2d", iters, px, [&] { CK(cudaMemcpyAsync(dIn, pinIn, inBytes, cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s)); });
    timeit("d2h", iters, px, [&] { CK(cudaMemcpyAsync(pinOut, dOut, outBytes, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); });
    timeit("kernel", iters, px, [&] {
        if (goofy_b200_encode_device(GOOFY_B200_DXT1, dOut, dIn, W, H, W * 4, s) != 0) std::exit(3);
        CK(cudaStreamSynchronize(s));
    });
    timeit("bare", iters, px, [&] {
        CK(cudaMemcpyAsync(dIn, pinIn, inBytes, cudaMemcpyHostToDevice, s));
        if (goofy_b200_encode_device(GOOFY_B200_DXT1, dOut, dIn, W, H, W * 4, s) != 0) std::exit(3);
        CK(cudaMemcpyAsync(pinOut, dOut, outBytes, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    });
    std::memcpy(want, pinOut, outBytes);
    timeit("lib pinned", iters, px, [&] { if (goofy_b200_compress_dxt1(pinOut, pinIn, W, H, W * 4) != 0) std::exit(3); });
    const bool okPinned = std::memcmp(want, pinOut, outBytes) == 0;
    timeit("lib pageable", iters, px, [&] { if (goofy_b200_compress_dxt1(pgOut, pgIn, W, H, W * 4) != 0) std::exit(3); });
    const bool okPageable = std::memcmp(want, pgOut, outBytes) == 0;
    std::printf("same bytes as the bare sequence: pinned %s, pageable %s\n", okPinned ? "yes" : "NO", okPageable ? "yes" : "NO");
    return okPinned && okPageable ? 0 : 5;
}
