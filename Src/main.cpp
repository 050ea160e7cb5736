// Src/main.cpp -- benchmark / parity harness for the B200 encoders (C++ host side of the path).
//
// Counterpart of the reference harness (Src/main.cpp of the reference): it drives the encoders
// through the same function-pointer shape (CompressFunc_t, reference Src/main.cpp:644), keeps its
// measurement definitions -- best-of-N timing of the host call (:653-664, :702-713),
// MP/s = w*h / seconds / 1e6 (:953-958), RGB-PSNR with the 768 peak (:444,:466) -- and adds what
// the reference does not have: device-resident batches, a multi-GPU shard scheduler (one host
// thread and one stream per device, static partition by texture, no collectives) and CUDA-event
// timing.  With --images it is the reference's per-image loop (runTest, Src/main.cpp:836-986, and the image list
// :991-1039): PNG ingest under the reference loader's contract (include/goofy_png.h), best-of-N timing of the
// drop-in host call, device-resident timing, RGB-PSNR through the GPU decoder, a CSV row per encoder and image in the
// reference's column order, and -- when the compiled reference is at hand (oracle/_ref/libgoofy_ref.so, dlopen'ed,
// never linked) -- the SSE2 CPU path timed in the same run, single thread as shipped and row-parallel over T threads,
// with the B200 bytes compared against it.  Competitor encoders, the RGB565 baseline and the TGA writers of the
// reference harness are out of scope.
//
// CUDA C++ (the synthetic-texture generator is a kernel): built by Src/Makefile with
//   nvcc -x cu -gencode arch=compute_100a,code=sm_100a -Iinclude Src/main.cpp -Lgoofy_b200 -lgoofy_b200
//
//   goofy_bench [--codec dxt1|etc1|both] [--size 8192] [--textures 4] [--gpus N] [--iters 20]
//               [--stride-pad 0] [--host-iters 3] [--rgb24]                        synthetic textures
//               (--rgb24: every pass also from packed RGB8 copies of the textures, bytes compared)
//   goofy_bench --images DIR [--list a,b,c] [--host-iters 128] [--iters 200] [--csv FILE] [--cpu-ref LIB.so] [--save-dir DIR]
//                                                                                  the reference's image list
#include <cuda_runtime.h>
#include <dirent.h>
#include <dlfcn.h>

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

#include "goofy_containers.h"  // DDS / KTX writers (saveDds / saveKtx of the reference harness)
#include "goofy_png.h"         // PNG ingest under the reference loader's contract
#include "goofy_tc.h"          // include/goofy_tc.h: goofy::compressDXT1/ETC1 + goofy::b200::*

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

// RGBA8 rows -> packed RGB8 rows (the input of goofy::b200::encodeRgb24): harness data preparation only
__global__ void strip_alpha_kernel(uint8_t* dst, const uint8_t* src, uint32_t width, uint32_t height, uint32_t srcStride)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= width || y >= height) return;
    const uchar4 p = *reinterpret_cast<const uchar4*>(src + (size_t)y * srcStride + (size_t)x * 4);
    uint8_t* d = dst + ((size_t)y * width + x) * 3;
    d[0] = p.x;
    d[1] = p.y;
    d[2] = p.z;
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
    // image-list mode
    std::string imageDir, list, csv, cpuRef, saveDir;
    bool hostItersGiven = false, itersGiven = false;
    bool rgb24 = false;   // also run every pass from packed RGB8 copies of the textures
};

struct Shard {
    int device = 0;
    uint32_t first = 0, count = 0;  // texture indices
    uint8_t* src = nullptr;
    uint8_t* dst[2] = {nullptr, nullptr};
    uint64_t* sse = nullptr;
    cudaStream_t stream = nullptr;
    float ms[3] = {0, 0, 0};        // dxt1, etc1, dual (device time of `iters` passes)
    float msRgb[3] = {0, 0, 0};     // the same from packed RGB8 input (--rgb24)
    bool rgbSame = true;            // ... and its bytes equal the RGBA path's
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
        else if (a == "--textures") o.images = (uint32_t)std::atoi(next());
        else if (a == "--images") {   // a number: texture count of the synthetic mode (round-1 spelling); otherwise a directory of PNGs
            const char* v = next();
            if (*v && std::strspn(v, "0123456789") == std::strlen(v)) o.images = (uint32_t)std::atoi(v);
            else o.imageDir = v;
        }
        else if (a == "--list") o.list = next();
        else if (a == "--csv") o.csv = next();
        else if (a == "--cpu-ref") o.cpuRef = next();
        else if (a == "--save-dir") o.saveDir = next();
        else if (a == "--gpus") o.gpus = std::atoi(next());
        else if (a == "--iters") { o.iters = (uint32_t)std::atoi(next()); o.itersGiven = true; }
        else if (a == "--stride-pad") o.stridePad = (uint32_t)std::atoi(next());
        else if (a == "--rgb24") o.rgb24 = true;
        else if (a == "--host-iters") { o.hostIters = (uint32_t)std::atoi(next()); o.hostItersGiven = true; }
        else {
            std::fprintf(stderr, "usage: goofy_bench [--codec dxt1|etc1|both] [--size N] [--textures K] [--gpus G] [--iters I] "
                                 "[--stride-pad BYTES] [--host-iters I] [--rgb24]\n"
                                 "       goofy_bench --images DIR [--list a,b,c] [--host-iters 128] [--iters 200] [--csv FILE] [--cpu-ref LIB.so] "
                                 "[--save-dir DIR]\n");
            std::exit(1);
        }
    }
    return o;
}


// ================================================================================ image-list mode
// The compiled reference (oracle/_ref/libgoofy_ref.so: goofy::compress* behind oracle/ref_shim.cpp), loaded at run time
// if it is there.  It is the timed CPU baseline and the byte-for-byte checker of this mode; the encoders never see it.
struct CpuReference {
    typedef int (*Fn)(unsigned char*, const unsigned char*, unsigned, unsigned, unsigned);
    typedef int (*FnMt)(int, unsigned char*, const unsigned char*, unsigned, unsigned, unsigned, int);
    void* handle = nullptr;
    Fn single[2] = {nullptr, nullptr};
    FnMt parallel = nullptr;
    unsigned threads = 1;
    bool open(const std::string& path)
    {
        handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!handle) return false;
        single[0] = (Fn)dlsym(handle, "ref_goofy_compress_dxt1");
        single[1] = (Fn)dlsym(handle, "ref_goofy_compress_etc1");
        parallel = (FnMt)dlsym(handle, "ref_goofy_compress_mt");
        unsigned (*hw)() = (unsigned (*)())dlsym(handle, "ref_hardware_threads");
        if (hw) threads = std::max(1u, hw());
        return single[0] && single[1] && parallel;
    }
};

template <typename F>
static double bestOfMicros(uint32_t iters, F fn)   // the reference's protocol: best of N calls (Src/main.cpp:653-664)
{
    double best = 1e30;
    for (uint32_t k = 0; k < iters; ++k) {
        const auto a = std::chrono::steady_clock::now();
        fn();
        best = std::min(best, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count());
    }
    return best;
}

static std::string exeDir(const char* argv0)
{
    std::string p = argv0;
    const size_t slash = p.rfind('/');
    return slash == std::string::npos ? "." : p.substr(0, slash);
}

static int runImageList(const Options& opt, const char* argv0)
{
    // the list: --list a,b,c, else every *.png of the directory in name order
    std::vector<std::string> names;
    if (!opt.list.empty()) {
        size_t pos = 0;
        while (pos <= opt.list.size()) {
            const size_t comma = opt.list.find(',', pos);
            const std::string n = opt.list.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
            if (!n.empty()) names.push_back(n);
            if (comma == std::string::npos) break;
            pos = comma + 1;
        }
    } else {
        DIR* d = opendir(opt.imageDir.c_str());
        if (!d) { std::fprintf(stderr, "cannot open directory %s\n", opt.imageDir.c_str()); return 1; }
        while (dirent* e = readdir(d)) {
            const std::string n = e->d_name;
            if (n.size() > 4 && n.substr(n.size() - 4) == ".png") names.push_back(n.substr(0, n.size() - 4));
        }
        closedir(d);
        std::sort(names.begin(), names.end());
    }
    CpuReference cpu;
    const std::string refPath = opt.cpuRef.empty() ? exeDir(argv0) + "/../oracle/_ref/libgoofy_ref.so" : opt.cpuRef;
    const bool haveCpu = cpu.open(refPath);
    const uint32_t hostIters = opt.hostItersGiven ? std::max(1u, opt.hostIters) : 128u;   // kNumberOfIterations, Src/main.cpp:43
    const uint32_t devIters = opt.itersGiven ? std::max(1u, opt.iters) : 200u;
    FILE* csv = opt.csv.empty() ? nullptr : std::fopen(opt.csv.c_str(), "w");
    auto row = [&](const char* image, const char* encoder, const char* format, double pixels, double micros, double psnr768, double psnrText,
                   const char* exact) {
        // the reference's column order (Src/main.cpp:948-962): image;encoder;format;pixels;microseconds;MP/s;quality...
        char line[512];
        std::snprintf(line, sizeof(line), "%s;%s;%s;%3.0f;%3.2f;%3.5f;%3.5f;%3.5f;%s\n", image, encoder, format, pixels, micros,
                      micros > 0 ? pixels / micros : 0.0, psnr768, psnrText, exact);
        std::fputs(line, stdout);
        if (csv) std::fputs(line, csv);
    };
    std::printf("image;encoder;format;pixels;best_us;MP/s;psnrRGB(768 peak);psnr(textbook);bit_exact_vs_sse2\n");
    if (csv) std::fprintf(csv, "image;encoder;format;pixels;best_us;MP/s;psnrRGB(768 peak);psnr(textbook);bit_exact_vs_sse2\n");

    CK(cudaSetDevice(0));
    cudaStream_t stream;
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    uint64_t* dSse;
    CK(cudaMalloc(&dSse, 3 * sizeof(uint64_t)));
    const char* formats[2] = {"DXT1", "ETC1"};
    const CompressFunc_t fns[2] = {goofy::compressDXT1, goofy::compressETC1};
    // sums for the closing summary: [codec][0 host call, 1 device, 2 cpu 1T, 3 cpu NT] MP/s, PSNR
    double sumMps[2][4] = {}, sumPsnr[2] = {}, sumPsnrKodak[2] = {};
    uint32_t loaded = 0, kodak = 0, skipped = 0, mismatches = 0;
    for (const std::string& name : names) {
        const std::string path = opt.imageDir + "/" + name + ".png";
        goofy::png::Image im = goofy::png::load(path.c_str());
        if (!im.error.empty()) {
            std::fprintf(stderr, "%s: %s\n", path.c_str(), im.error.c_str());   // the reference harness skips such images too
            ++skipped;
            continue;
        }
        const uint32_t W = im.width, H = im.height, stride = W * 4;
        const size_t inBytes = (size_t)stride * H, outBytes = (size_t)W * H / 2;
        const double px = (double)W * H;
        unsigned char* result = (unsigned char*)std::aligned_alloc(64, (outBytes + 63) / 64 * 64);
        unsigned char* cpuResult = (unsigned char*)std::aligned_alloc(64, (outBytes + 63) / 64 * 64);
        uint8_t *dSrc, *dDst;
        CK(cudaMalloc(&dSrc, inBytes));
        CK(cudaMalloc(&dDst, outBytes));
        CK(cudaMemcpy(dSrc, im.rgba, inBytes, cudaMemcpyHostToDevice));
        const bool isKodak = name.compare(0, 5, "kodim") == 0;
        for (int c = 0; c < 2; ++c) {
            if ((c == 0 && opt.codec == "etc1") || (c == 1 && opt.codec == "dxt1")) continue;
            // the drop-in host call on ordinary (malloc'ed, 64-byte aligned) buffers, as the reference harness times it
            GK(fns[c](result, im.rgba, W, H, stride));
            const double hostUs = bestOfMicros(hostIters, [&] { GK(fns[c](result, im.rgba, W, H, stride)); });
            // device-resident: back-to-back launches between two events
            for (int k = 0; k < 5; ++k) GK(goofy::b200::encode((goofy::b200::Codec)c, dDst, dSrc, W, H, stride, stream));
            float ms = 0, bestMs = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0, stream));
                for (uint32_t k = 0; k < devIters; ++k) GK(goofy::b200::encode((goofy::b200::Codec)c, dDst, dSrc, W, H, stride, stream));
                CK(cudaEventRecord(e1, stream));
                CK(cudaStreamSynchronize(stream));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                bestMs = std::min(bestMs, ms);
            }
            const double devUs = bestMs * 1e3 / devIters;
            // quality: decode + squared error on the device (goofy::b200::blockSse), the harness's psnrRGB formula (:444,:466)
            uint64_t sse[3];
            CK(cudaMemsetAsync(dSse, 0, sizeof(sse), stream));
            GK(goofy::b200::blockSse((goofy::b200::Codec)c, dDst, dSrc, W, H, stride, dSse, stream));
            CK(cudaMemcpyAsync(sse, dSse, sizeof(sse), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            const double mseSum = (double)(sse[0] + sse[1] + sse[2]) / px;
            const double psnr768 = mseSum > 0 ? 10.0 * std::log10(768.0 * 768.0 / mseSum) : 999.0;
            const double psnrText = mseSum > 0 ? 10.0 * std::log10(255.0 * 255.0 / (mseSum / 3.0)) : 999.0;
            const char* exact = "n/a";
            double cpu1 = 0, cpuN = 0;
            if (haveCpu) {
                cpu.single[c](cpuResult, im.rgba, W, H, stride);
                exact = std::memcmp(cpuResult, result, outBytes) == 0 ? "yes" : "NO";
                if (exact[0] == 'N') ++mismatches;
                cpu1 = bestOfMicros(hostIters, [&] { cpu.single[c](cpuResult, im.rgba, W, H, stride); });
                cpuN = bestOfMicros(std::max(8u, hostIters / 4), [&] { cpu.parallel(c, cpuResult, im.rgba, W, H, stride, (int)cpu.threads); });
            }
            row(name.c_str(), "b200_goofy (host call)", formats[c], px, hostUs, psnr768, psnrText, exact);
            row(name.c_str(), "b200_goofy (device-resident)", formats[c], px, devUs, psnr768, psnrText, exact);
            if (haveCpu) {
                row(name.c_str(), "simd_goofy (cpu, 1 thread)", formats[c], px, cpu1, psnr768, psnrText, "reference");
                char enc[64];
                std::snprintf(enc, sizeof(enc), "simd_goofy (cpu, %u threads)", cpu.threads);
                row(name.c_str(), enc, formats[c], px, cpuN, psnr768, psnrText, "reference");
            }
            sumMps[c][0] += px / hostUs;
            sumMps[c][1] += px / devUs;
            if (haveCpu) { sumMps[c][2] += px / cpu1; sumMps[c][3] += px / cpuN; }
            sumPsnr[c] += psnr768;
            if (isKodak) sumPsnrKodak[c] += psnr768;
            if (!opt.saveDir.empty()) {   // saveDds / saveKtx of the reference harness (:154-220)
                const std::string out = opt.saveDir + "/" + name + (c == 0 ? ".dds" : ".ktx");
                const bool ok = c == 0 ? goofy::containers::writeDdsDxt1(out.c_str(), result, W, H) : goofy::containers::writeKtxEtc1(out.c_str(), result, W, H);
                if (!ok) std::fprintf(stderr, "cannot write %s\n", out.c_str());
            }
        }
        ++loaded;
        if (isKodak) ++kodak;
        CK(cudaFree(dSrc));
        CK(cudaFree(dDst));
        std::free(result);
        std::free(cpuResult);
        goofy::png::freeImage(im);
    }
    if (csv) std::fclose(csv);
    std::printf("{\"harness\": \"Src/main.cpp --images\", \"images\": %u, \"skipped\": %u, \"host_iters\": %u, \"device_iters\": %u, \"cpu_reference\": %s, "
                "\"cpu_threads\": %u, \"mismatches_vs_sse2\": %u",
                loaded, skipped, hostIters, devIters, haveCpu ? "true" : "false", haveCpu ? cpu.threads : 0u, mismatches);
    const char* names2[2] = {"dxt1", "etc1"};
    for (int c = 0; c < 2; ++c) {
        if (loaded == 0 || sumMps[c][0] == 0) continue;
        std::printf(", \"%s\": {\"mean_mp_per_s_host_call\": %.1f, \"mean_mp_per_s_device\": %.1f, \"mean_mp_per_s_cpu_1_thread\": %.1f, "
                    "\"mean_mp_per_s_cpu_all_threads\": %.1f, \"mean_psnr_rgb768\": %.3f, \"kodak_images\": %u, \"kodak_mean_psnr_rgb768\": %.3f}",
                    names2[c], sumMps[c][0] / loaded, sumMps[c][1] / loaded, sumMps[c][2] / loaded, sumMps[c][3] / loaded, sumPsnr[c] / loaded, kodak,
                    kodak ? sumPsnrKodak[c] / kodak : 0.0);
    }
    std::printf("}\n");
    return mismatches ? 5 : (loaded ? 0 : 6);
}

int main(int argc, char** argv)
{
    const Options opt = parse(argc, argv);
    const int visible = goofy::b200::deviceCount();
    if (visible <= 0) {
        std::fprintf(stderr, "no CUDA device: the encoders have no CPU fallback\n");
        return 4;
    }
    if (!opt.imageDir.empty()) return runImageList(opt, argv[0]);
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

        // ---- the same textures as packed RGB8 (3 bytes per pixel, tight rows) through goofy::b200::encodeRgb24
        uint8_t* rgb = nullptr;
        uint8_t* rgbOut[2] = {nullptr, nullptr};
        const size_t rgbBytes = (size_t)W * H * 3;
        if (opt.rgb24 && s.count) {
            CK(cudaMalloc(&rgb, s.count * rgbBytes));
            CK(cudaMalloc(&rgbOut[0], s.count * outBytes));
            CK(cudaMalloc(&rgbOut[1], s.count * outBytes));
            for (uint32_t i = 0; i < s.count; ++i)
                strip_alpha_kernel<<<dim3((W + 255) / 256, H), 256, 0, s.stream>>>(rgb + (size_t)i * rgbBytes, s.src + (size_t)i * imgBytes, W, H, stride);
            CK(cudaGetLastError());
            for (int mode = 0; mode < 3; ++mode) {
                if (!(mode < 2 ? doCodec[mode] : doDual)) continue;
                auto pass = [&]() {
                    GK(goofy::b200::encodeRgb24(mode, rgbOut[mode == 1 ? 1 : 0], mode == 2 ? rgbOut[1] : nullptr, rgb, W, H, W * 3, rgbBytes, outBytes,
                                                s.count, s.stream));
                };
                for (int k = 0; k < 3; ++k) pass();
                CK(cudaEventRecord(e0, s.stream));
                for (uint32_t k = 0; k < opt.iters; ++k) pass();
                CK(cudaEventRecord(e1, s.stream));
                CK(cudaStreamSynchronize(s.stream));
                CK(cudaEventElapsedTime(&s.msRgb[mode], e0, e1));
            }
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
            if (rgb) {   // packed-RGB input must give the same bytes
                GK(goofy::b200::encodeRgb24(c, rgbOut[c], nullptr, rgb, W, H, W * 3, rgbBytes, outBytes, s.count, s.stream));
                for (uint32_t i = 0; i < s.count; ++i) {
                    CK(cudaMemcpyAsync(host.data(), rgbOut[c] + (size_t)i * outBytes, outBytes, cudaMemcpyDeviceToHost, s.stream));
                    CK(cudaStreamSynchronize(s.stream));
                    s.rgbSame = s.rgbSame && fnv1a(host.data(), outBytes) == s.hash[c][i];
                }
            }
        }
        CK(cudaStreamSynchronize(s.stream));
        if (rgb) {
            CK(cudaFree(rgb));
            CK(cudaFree(rgbOut[0]));
            CK(cudaFree(rgbOut[1]));
        }
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
    if (opt.rgb24) {
        const double rgbBytesPerPx[3] = {3.5, 3.5, 4.0};
        bool same = true;
        for (auto& s : shards) same = same && s.rgbSame;
        std::printf(", \"rgb24_input\": {\"equals_rgba_path\": %s", same ? "true" : "false");
        for (int mode = 0; mode < 3; ++mode) {
            if (!(mode < 2 ? doCodec[mode] : doDual)) continue;
            float msMax = 0;
            for (auto& s : shards) msMax = std::max(msMax, s.msRgb[mode]);
            const double mps = totalPx * opt.iters / (msMax * 1e-3) / 1e6;
            std::printf(", \"%s\": {\"mp_per_s\": %.1f, \"gb_per_s_per_gpu\": %.1f}", names[mode], mps, mps * 1e6 * rgbBytesPerPx[mode] / 1e9 / G);
        }
        std::printf("}");
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
            uint64_t up0 = 0, up1 = 0;   // bytes the host path sent over the link for one more call (alpha-stripped strips send 3/4)
            goofy_b200_host_link_stats(&up0, nullptr, nullptr, nullptr, nullptr);
            GK(fns[c](hdst, hsrc, W, H, stride));
            goofy_b200_host_link_stats(&up1, nullptr, nullptr, nullptr, nullptr);
            std::printf(", \"host_api_%s\": {\"mp_per_s\": %.1f, \"best_us\": %.1f, \"equals_device_path\": %s, \"h2d_bytes_last_call\": %llu, "
                        "\"h2d_bytes_logical\": %llu}",
                        names[c], ((double)W * H / (best / 1e6)) / 1e6, best, same ? "true" : "false", (unsigned long long)(up1 - up0),
                        (unsigned long long)W * H * 4ull);
        }
        CK(cudaFreeHost(hsrc));
        CK(cudaFreeHost(hdst));
    }
    std::printf("}\n");
    return identical ? 0 : 5;
}
