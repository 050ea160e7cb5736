// Src/main.cpp -- benchmark / parity harness for the B200 encoders (C++ host side of the path).
//
// Counterpart of the reference harness (Src/main.cpp of the reference): it drives the encoders
// through the same function-pointer shape (CompressFunc_t, reference Src/main.cpp:644), keeps its
// measurement definitions -- best-of-N timing of the host call (:653-664, :702-713),
// MP/s = w*h / seconds / 1e6 (:953-958), RGB-PSNR with the 768 peak (:444,:466) -- and adds what
// the reference does not have: device-resident batches, a multi-GPU shard scheduler (one host
// thread and one stream per device, static partition by texture, no collectives) and CUDA-event
// timing.  File writers, competitor encoders and the PNG loader of the reference harness are
// out of scope.
//
// CUDA C++ (the synthetic-texture generator is a kernel): built by Src/Makefile with
//   nvcc -x cu -gencode arch=compute_100a,code=sm_100a -Iinclude Src/main.cpp -Lgoofy_b200 -lgoofy_b200
//
//   goofy_bench [--codec dxt1|etc1|both] [--size 8192] [--images 4] [--gpus N] [--iters 20]
//               [--stride-pad 0] [--host-iters 3]
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "goofy_tc.h"  // include/goofy_tc.h: goofy::compressDXT1/ETC1 + goofy::b200::*

typedef int (*CompressFunc_t)(unsigned char* result, const unsigned char* input, unsigned int width, unsigned int height,
                              unsigned int stride);

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            std::fprintf(stderr, "%s:%d CUDA error %s\n", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            std::exit(2);                                                                     \
        }                                                                                     \
    } while (0)
#define GK(call)                                                                              \
    do {                                                                                      \
        int r__ = (call);                                                                     \
        if (r__ != 0) {                                                                       \
            std::fprintf(stderr, "%s:%d goofy_b200 error %d: %s\n", __FILE__, __LINE__, r__, goofy::b200::errorString(r__)); \
            std::exit(3);                                                                     \
        }                                                                                     \
    } while (0)

// Counter-based synthetic texture (family S1 of SURVEY.md 8(d): smooth gradient + 4-bit noise), any
// pixel computable independently: z = splitmix64(seed + index * golden); byte = ((x+y)/8 + (z&15) + 20c) & 255.
__device__ __host__ inline uint64_t splitmix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void fill_texture_kernel(uint8_t* dst, uint32_t width, uint32_t height, uint32_t stride, uint64_t seed)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= width || y >= height) return;
    const uint64_t z = splitmix64(seed + ((uint64_t)y * width + x) * 0x9E3779B97F4A7C15ull);
    const uint32_t base = (x + y) / 8u;
    uchar4 p;
    p.x = (uint8_t)(base + (z & 15u));
    p.y = (uint8_t)(base + ((z >> 4) & 15u) + 20u);
    p.z = (uint8_t)(base + ((z >> 8) & 15u) + 40u);
    p.w = (uint8_t)(base + ((z >> 12) & 15u) + 60u);
    *reinterpret_cast<uchar4*>(dst + (size_t)y * stride + (size_t)x * 4) = p;
}

static uint64_t fnv1a(const uint8_t* p, size_t n)
{
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}

struct Options {
    std::string codec = "both";
    uint32_t size = 8192, images = 4, iters = 20, stridePad = 0, hostIters = 3;
    int gpus = 1;
};

struct Shard {
    int device = 0;
    uint32_t first = 0, count = 0;  // texture indices
    uint8_t* src = nullptr;
    uint8_t* dst[2] = {nullptr, nullptr};
    uint64_t* sse = nullptr;
    cudaStream_t stream = nullptr;
    float ms[3] = {0, 0, 0};        // dxt1, etc1, dual (device time of `iters` passes)
    uint64_t sseHost[2][3] = {};
    std::vector<uint64_t> hash[2];
};

static Options parse(int argc, char** argv)
{
    Options o;
    for (int i = 1; i < argc; ++i) {
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        const std::string a = argv[i];
        if (a == "--codec") o.codec = next();
        else if (a == "--size") o.size = (uint32_t)std::atoi(next());
        else if (a == "--images") o.images = (uint32_t)std::atoi(next());
        else if (a == "--gpus") o.gpus = std::atoi(next());
        else if (a == "--iters") o.iters = (uint32_t)std::atoi(next());
        else if (a == "--stride-pad") o.stridePad = (uint32_t)std::atoi(next());
        else if (a == "--host-iters") o.hostIters = (uint32_t)std::atoi(next());
        else {
            std::fprintf(stderr, "usage: goofy_bench [--codec dxt1|etc1|both] [--size N] [--images K] [--gpus G] [--iters I] "
                                 "[--stride-pad BYTES] [--host-iters I]\n");
            std::exit(1);
        }
    }
    return o;
}

int main(int argc, char** argv)
{
    const Options opt = parse(argc, argv);
    const int visible = goofy::b200::deviceCount();
    if (visible <= 0) {
        std::fprintf(stderr, "no CUDA device: the encoders have no CPU fallback\n");
        return 4;
    }
    const int G = std::min(opt.gpus, visible);
    const uint32_t W = opt.size, H = opt.size, stride = W * 4 + opt.stridePad;
    const size_t imgBytes = (size_t)stride * H, outBytes = (size_t)W * H / 2;
    const bool doCodec[2] = {opt.codec != "etc1", opt.codec != "dxt1"};
    const bool doDual = opt.codec == "both";

    // ---- static partition of the texture batch: contiguous ranges, sizes differ by at most one
    std::vector<Shard> shards((size_t)G);
    for (int g = 0; g < G; ++g) {
        shards[g].device = g;
        shards[g].first = (uint32_t)((uint64_t)opt.images * g / G);
        shards[g].count = (uint32_t)((uint64_t)opt.images * (g + 1) / G) - shards[g].first;
    }

    std::atomic<int> arrived{0};
    auto barrier = [&](int phase) {  // host barrier between phases
        arrived.fetch_add(1);
        while (arrived.load() < G * phase) std::this_thread::yield();
    };
    std::vector<double> wall(3, 0.0);
    std::chrono::steady_clock::time_point t0[3], t1[3];

    auto worker = [&](int g) {
        Shard& s = shards[g];
        CK(cudaSetDevice(s.device));
        CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        const size_t n = s.count ? s.count : 1;
        CK(cudaMalloc(&s.src, n * imgBytes));
        CK(cudaMalloc(&s.dst[0], n * outBytes));
        CK(cudaMalloc(&s.dst[1], n * outBytes));
        CK(cudaMalloc(&s.sse, 6 * sizeof(uint64_t)));
        CK(cudaMemsetAsync(s.src, 0xAB, n * imgBytes, s.stream));  // pad bytes must be ignored
        for (uint32_t i = 0; i < s.count; ++i)
            fill_texture_kernel<<<dim3((W + 255) / 256, H), 256, 0, s.stream>>>(s.src + (size_t)i * imgBytes, W, H, stride,
                                                                              0x9E3779B97F4A7C15ull + s.first + i);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(s.stream));

        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        int phase = 0;
        for (int mode = 0; mode < 3; ++mode) {
            const bool run = mode < 2 ? doCodec[mode] : doDual;
            auto pass = [&]() {
                if (!s.count) return;
                if (mode < 2)
                    GK(goofy::b200::encodeBatch((goofy::b200::Codec)mode, s.dst[mode], s.src, W, H, stride, imgBytes, outBytes,
                                                s.count, s.stream));
                else
                    GK(goofy::b200::encodeDual(s.dst[0], s.dst[1], s.src, W, H, stride, imgBytes, outBytes, s.count, s.stream));
            };
            if (run) {
                for (int k = 0; k < 3; ++k) pass();
                CK(cudaStreamSynchronize(s.stream));
            }
            barrier(++phase);
            if (g == 0) t0[mode] = std::chrono::steady_clock::now();
            if (run) {
                CK(cudaEventRecord(e0, s.stream));
                for (uint32_t k = 0; k < opt.iters; ++k) pass();
                CK(cudaEventRecord(e1, s.stream));
                CK(cudaStreamSynchronize(s.stream));
                CK(cudaEventElapsedTime(&s.ms[mode], e0, e1));
            }
            barrier(++phase);
            if (g == 0) t1[mode] = std::chrono::steady_clock::now();
        }

        // ---- quality and identity of the results, computed where the data lives
        std::vector<uint8_t> host(outBytes);
        for (int c = 0; c < 2; ++c) {
            if (!doCodec[c] || !s.count) continue;
            GK(goofy::b200::encodeBatch((goofy::b200::Codec)c, s.dst[c], s.src, W, H, stride, imgBytes, outBytes, s.count, s.stream));
            CK(cudaMemsetAsync(s.sse + 3 * c, 0, 3 * sizeof(uint64_t), s.stream));
            for (uint32_t i = 0; i < s.count; ++i)
                GK(goofy_b200_block_sse_device(c, s.dst[c] + (size_t)i * outBytes, s.src + (size_t)i * imgBytes, W, H, stride,
                                               s.sse + 3 * c, s.stream));
            CK(cudaMemcpyAsync(s.sseHost[c], s.sse + 3 * c, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s.stream));
            for (uint32_t i = 0; i < s.count; ++i) {
                CK(cudaMemcpyAsync(host.data(), s.dst[c] + (size_t)i * outBytes, outBytes, cudaMemcpyDeviceToHost, s.stream));
                CK(cudaStreamSynchronize(s.stream));
                s.hash[c].push_back(fnv1a(host.data(), outBytes));
            }
        }
        CK(cudaStreamSynchronize(s.stream));
    };

    std::vector<std::thread> pool;
    for (int g = 0; g < G; ++g) pool.emplace_back(worker, g);
    for (auto& t : pool) t.join();

    const double totalPx = (double)W * H * opt.images;
    const char* names[3] = {"dxt1", "etc1", "dual"};
    const double bytesPerPx[3] = {4.5, 4.5, 5.0};
    std::printf("{\"harness\": \"Src/main.cpp\", \"size\": %u, \"images\": %u, \"gpus\": %d, \"stride\": %u, \"iters\": %u", W, opt.images,
                G, stride, opt.iters);
    for (int mode = 0; mode < 3; ++mode) {
        if (!(mode < 2 ? doCodec[mode] : doDual)) continue;
        float msMax = 0;
        for (auto& s : shards) msMax = std::max(msMax, s.ms[mode]);
        const double wallS = std::chrono::duration<double>(t1[mode] - t0[mode]).count();
        const double mps = totalPx * opt.iters / (msMax * 1e-3) / 1e6;
        std::printf(", \"%s\": {\"mp_per_s\": %.1f, \"device_ms_max\": %.4f, \"wall_ms\": %.4f, \"gb_per_s_per_gpu\": %.1f, "
                    "\"frac_of_8TBs\": %.4f}",
                    names[mode], mps, msMax, wallS * 1e3, mps * 1e6 * bytesPerPx[mode] / 1e9 / G, mps * 1e6 * bytesPerPx[mode] / 1e9 / G / 8000.0);
    }
    for (int c = 0; c < 2; ++c) {
        if (!doCodec[c]) continue;
        uint64_t sse = 0;
        for (auto& s : shards) sse += s.sseHost[c][0] + s.sseHost[c][1] + s.sseHost[c][2];
        const double mse = (double)sse / totalPx;
        std::printf(", \"psnr_rgb768_%s\": %.3f, \"psnr_textbook_%s\": %.3f", names[c], 10.0 * std::log10(768.0 * 768.0 / mse), names[c],
                    10.0 * std::log10(255.0 * 255.0 / (mse / 3.0)));
    }

    // ---- N-GPU output must equal the 1-GPU output byte for byte: re-encode every texture on device 0
    bool identical = true;
    if (G > 1) {
        CK(cudaSetDevice(0));
        uint8_t *src0, *dst0;
        CK(cudaMalloc(&src0, imgBytes));
        CK(cudaMalloc(&dst0, outBytes));
        std::vector<uint8_t> host(outBytes);
        for (auto& s : shards)
            for (uint32_t i = 0; i < s.count; ++i) {
                CK(cudaMemset(src0, 0xAB, imgBytes));
                fill_texture_kernel<<<dim3((W + 255) / 256, H), 256>>>(src0, W, H, stride, 0x9E3779B97F4A7C15ull + s.first + i);
                for (int c = 0; c < 2; ++c) {
                    if (!doCodec[c]) continue;
                    GK(goofy::b200::encode((goofy::b200::Codec)c, dst0, src0, W, H, stride, nullptr));
                    CK(cudaMemcpy(host.data(), dst0, outBytes, cudaMemcpyDeviceToHost));
                    identical = identical && fnv1a(host.data(), outBytes) == s.hash[c][i];
                }
            }
        CK(cudaFree(src0));
        CK(cudaFree(dst0));
    }
    std::printf(", \"multi_gpu_equals_single_gpu\": %s", identical ? "true" : "false");

    // ---- the drop-in host call, timed the way the reference harness times it (best of N, Src/main.cpp:653-664)
    if (opt.hostIters) {
        CK(cudaSetDevice(0));
        uint8_t *hsrc, *hdst;
        CK(cudaHostAlloc(&hsrc, imgBytes, cudaHostAllocDefault));
        CK(cudaHostAlloc(&hdst, outBytes, cudaHostAllocDefault));
        CK(cudaMemcpy(hsrc, shards[0].src, imgBytes, cudaMemcpyDeviceToHost));
        const CompressFunc_t fns[2] = {goofy::compressDXT1, goofy::compressETC1};
        for (int c = 0; c < 2; ++c) {
            if (!doCodec[c]) continue;
            double best = 1e30;
            for (uint32_t k = 0; k < opt.hostIters + 1; ++k) {
                const auto a = std::chrono::steady_clock::now();
                GK(fns[c](hdst, hsrc, W, H, stride));
                const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count();
                if (k) best = std::min(best, us);
            }
            const bool same = shards[0].count && fnv1a(hdst, outBytes) == shards[0].hash[c][0];
            std::printf(", \"host_api_%s\": {\"mp_per_s\": %.1f, \"best_us\": %.1f, \"equals_device_path\": %s}", names[c],
                        ((double)W * H / (best / 1e6)) / 1e6, best, same ? "true" : "false");
        }
        CK(cudaFreeHost(hsrc));
        CK(cudaFreeHost(hdst));
    }
    std::printf("}\n");
    return identical ? 0 : 5;
}
