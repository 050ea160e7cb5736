// shapebench.cu -- A/B of library builds / load layers / environment knobs on the NAMED launch shapes, in one
// process on one box, with no Python between launches (a 16384 x 2048 strip is a 21 us kernel: the host has to
// keep up).  Experiments only; the product is goofy_b200/libgoofy_b200.so behind include/goofy_b200.h.
//
//   shapebench [--iters 200] [--rounds 3] [--shapes strip,tex8192,batch1024,batch4x8192] [--modes dxt1,etc1,dual]
//              [--json out.json] name=lib.so[:KEY=VAL...][:path=N] ...
//
// Every variant gets its OWN copy of the library (copied to a temp file and dlopen'ed), so the environment
// overrides it reads on first use (GOOFY_B200_*) and its load path are private to it.  All variants must produce
// the same bytes (checked with a hash of the outputs against the first variant).
//   strip        16384 x 2048 RGBA8, stride 65792 (pad 0xAB)   -- one rank's share of BASELINE configs[4] at 8 GPUs
//   tex8192      one 8192^2 texture per launch, four textures rotated      -- configs[1]/[2], strict form
//   batch1024    512 x 1024^2 as one uniform batch launch                  -- one rank's share of configs[3] at 8 GPUs
//   batch4x8192  4 x 8192^2 in one launch                                  -- what bench.py times
//   strip1k      16384 x 1024 strip (16 ranks' worth: an even shorter launch)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            std::fprintf(stderr, "%s:%d CUDA error %s\n", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            std::exit(2);                                                                               \
        }                                                                                               \
    } while (0)

typedef int (*BatchFn)(int, void*, const void*, uint32_t, uint32_t, uint32_t, uint64_t, uint64_t, uint32_t, void*);
typedef int (*DualFn)(void*, void*, const void*, uint32_t, uint32_t, uint32_t, uint64_t, uint64_t, uint32_t, void*);
typedef int (*SetPathFn)(int);

struct Variant {
    std::string name;
    void* handle = nullptr;
    BatchFn batch = nullptr;
    DualFn dual = nullptr;
};

struct Shape {
    std::string name;
    uint32_t w, h, stride, images, rotate;   // `rotate` distinct inputs cycled launch by launch
    uint32_t pitchPad = 0;                   // extra bytes between the images of a batch (pitched batch)
};

__device__ __host__ inline uint64_t splitmix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// S1 of SURVEY.md 8(d): smooth gradient + 4-bit noise; pad bytes keep the 0xAB they were memset to
__global__ void fill_kernel(uint8_t* dst, uint32_t width, uint32_t height, uint32_t stride, uint64_t seed)
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

__global__ void hash_kernel(const uint64_t* p, size_t n, unsigned long long* out)
{
    unsigned long long h = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        h += splitmix64(p[i] + i * 0x9E3779B97F4A7C15ull);
    atomicAdd(out, h);
}

static std::vector<std::string> split(const std::string& s, char c)
{
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, c)) out.push_back(item);
    return out;
}

int main(int argc, char** argv)
{
    uint32_t iters = 60, rounds = 9;
    std::string shapesArg = "strip,tex8192,batch1024,batch4x8192", modesArg = "dxt1,etc1,dual", jsonPath;
    std::vector<std::string> specs;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--iters") iters = (uint32_t)std::atoi(next().c_str());
        else if (a == "--rounds") rounds = (uint32_t)std::atoi(next().c_str());
        else if (a == "--shapes") shapesArg = next();
        else if (a == "--modes") modesArg = next();
        else if (a == "--json") jsonPath = next();
        else specs.push_back(a);
    }
    if (specs.empty()) {
        std::fprintf(stderr, "usage: shapebench [--iters N] [--rounds R] [--shapes a,b] [--modes dxt1,etc1,dual] name=lib.so[:KEY=VAL][:path=N] ...\n");
        return 1;
    }
    const std::map<std::string, Shape> known = {
        {"strip", {"strip", 16384, 2048, 65792, 1, 2}},
        {"strip1k", {"strip1k", 16384, 1024, 65792, 1, 4}},
        {"tex8192", {"tex8192", 8192, 8192, 32768, 1, 4}},
        {"batch1024", {"batch1024", 1024, 1024, 4096, 512, 1}},
        {"batch4x8192", {"batch4x8192", 8192, 8192, 32768, 4, 1}},
        {"tex2048", {"tex2048", 2048, 2048, 8192, 1, 16}},
        {"batch1024p", {"batch1024p", 1024, 1024, 4096, 512, 1, 4096}},   // same batch, images 4 KiB apart: not one tall image
    };
    std::vector<Shape> shapes;
    for (auto& n : split(shapesArg, ',')) {
        auto it = known.find(n);
        if (it == known.end()) { std::fprintf(stderr, "unknown shape %s\n", n.c_str()); return 1; }
        shapes.push_back(it->second);
    }
    const std::vector<std::string> modes = split(modesArg, ',');

    CK(cudaSetDevice(0));
    cudaStream_t stream;
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));

    // buffers: the largest shape decides
    size_t maxIn = 0, maxOut = 0;
    for (auto& s : shapes) {
        maxIn = std::max(maxIn, ((size_t)s.stride * s.h + s.pitchPad) * s.images * s.rotate);
        maxOut = std::max(maxOut, (size_t)s.w * s.h / 2 * s.images);
    }
    uint8_t *src, *d1, *d2;
    unsigned long long* dHash;
    CK(cudaMalloc(&src, maxIn));
    CK(cudaMalloc(&d1, maxOut));
    CK(cudaMalloc(&d2, maxOut));
    CK(cudaMalloc(&dHash, 8));

    // load the variants
    std::vector<Variant> variants;
    for (auto& spec : specs) {
        const size_t eq = spec.find('=');
        if (eq == std::string::npos) { std::fprintf(stderr, "bad variant %s\n", spec.c_str()); return 1; }
        Variant v;
        v.name = spec.substr(0, eq);
        const std::vector<std::string> parts = split(spec.substr(eq + 1), ':');
        int path = -1;
        std::vector<std::string> envNames;
        for (size_t k = 1; k < parts.size(); ++k) {
            const size_t e = parts[k].find('=');
            if (e == std::string::npos) continue;
            const std::string key = parts[k].substr(0, e), val = parts[k].substr(e + 1);
            if (key == "path") path = std::atoi(val.c_str());
            else { setenv(key.c_str(), val.c_str(), 1); envNames.push_back(key); }
        }
        // private copy so that this variant's static state (environment overrides, load path) is its own
        char tmp[] = "/tmp/shapebench_XXXXXX";
        const int fd = mkstemp(tmp);
        if (fd < 0) { std::perror("mkstemp"); return 1; }
        close(fd);
        { std::ifstream in(parts[0], std::ios::binary); std::ofstream out(tmp, std::ios::binary); out << in.rdbuf(); if (!in.good() && !in.eof()) { std::fprintf(stderr, "cannot read %s\n", parts[0].c_str()); return 1; } }
        v.handle = dlopen(tmp, RTLD_NOW | RTLD_LOCAL);
        unlink(tmp);
        if (!v.handle) { std::fprintf(stderr, "dlopen %s: %s\n", parts[0].c_str(), dlerror()); return 1; }
        v.batch = (BatchFn)dlsym(v.handle, "goofy_b200_encode_batch_uniform_device");
        v.dual = (DualFn)dlsym(v.handle, "goofy_b200_encode_dual_device");
        SetPathFn setPath = (SetPathFn)dlsym(v.handle, "goofy_b200_set_load_path");
        if (!v.batch || !v.dual || !setPath) { std::fprintf(stderr, "%s: missing symbols\n", parts[0].c_str()); return 1; }
        if (path >= 0) setPath(path);
        // first use of every (shape, mode) under this variant's environment
        for (auto& s : shapes) {
            const uint64_t inPitch = (uint64_t)s.stride * s.h + s.pitchPad, outPitch = (uint64_t)s.w * s.h / 2;
            for (auto& m : modes) {
                int rc;
                if (m == "dual") rc = v.dual(d1, d2, src, s.w, s.h, s.stride, inPitch, outPitch, s.images, stream);
                else rc = v.batch(m == "etc1" ? 1 : 0, d1, src, s.w, s.h, s.stride, inPitch, outPitch, s.images, stream);
                if (rc != 0) { std::fprintf(stderr, "%s %s %s: rc %d\n", v.name.c_str(), s.name.c_str(), m.c_str(), rc); return 3; }
            }
        }
        CK(cudaStreamSynchronize(stream));
        for (auto& k : envNames) unsetenv(k.c_str());
        variants.push_back(v);
    }

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::ostringstream js;
    js << "{";
    bool allSame = true, firstShape = true;
    for (auto& s : shapes) {
        const size_t inPitch = (size_t)s.stride * s.h + s.pitchPad, outPitch = (size_t)s.w * s.h / 2;
        const size_t inBytes = inPitch * s.images;   // one rotation slot
        CK(cudaMemsetAsync(src, 0xAB, inBytes * s.rotate, stream));
        for (uint32_t r = 0; r < s.rotate; ++r)
            for (uint32_t i = 0; i < s.images; ++i)
                fill_kernel<<<dim3((s.w + 255) / 256, s.h), 256, 0, stream>>>(src + r * inBytes + i * inPitch, s.w, s.h, s.stride,
                                                                            1000u * r + (i % 61u) + 7u);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(stream));
        const double px = (double)s.w * s.h * s.images;
        std::printf("== %s: %u x %u, stride %u, %u image(s) per launch, %.1f MB algorithmic (4.5 B/px) per launch\n", s.name.c_str(), s.w, s.h,
                    s.stride, s.images, px * 4.5 / 1e6);
        js << (firstShape ? "" : ", ") << "\"" << s.name << "\": {";
        firstShape = false;
        bool firstMode = true;
        for (auto& m : modes) {
            const double bpp = m == "dual" ? 5.0 : 4.5;
            auto launch = [&](Variant& v, uint32_t k) {
                const uint8_t* in = src + (size_t)(k % s.rotate) * inBytes;
                int rc;
                if (m == "dual") rc = v.dual(d1, d2, in, s.w, s.h, s.stride, inPitch, outPitch, s.images, stream);
                else rc = v.batch(m == "etc1" ? 1 : 0, d1, in, s.w, s.h, s.stride, inPitch, outPitch, s.images, stream);
                if (rc != 0) { std::fprintf(stderr, "rc %d\n", rc); std::exit(3); }
            };
            // identity of the bytes across variants
            unsigned long long want[2] = {0, 0};
            for (size_t vi = 0; vi < variants.size(); ++vi) {
                CK(cudaMemsetAsync(d1, 0, outPitch * s.images, stream));
                CK(cudaMemsetAsync(d2, 0, outPitch * s.images, stream));
                launch(variants[vi], 0);
                unsigned long long got[2] = {0, 0};
                for (int o = 0; o < (m == "dual" ? 2 : 1); ++o) {
                    CK(cudaMemsetAsync(dHash, 0, 8, stream));
                    hash_kernel<<<1024, 256, 0, stream>>>((const uint64_t*)(o ? d2 : d1), outPitch * s.images / 8, dHash);
                    CK(cudaMemcpyAsync(&got[o], dHash, 8, cudaMemcpyDeviceToHost, stream));
                    CK(cudaStreamSynchronize(stream));
                }
                if (vi == 0) { want[0] = got[0]; want[1] = got[1]; }
                else if (got[0] != want[0] || got[1] != want[1]) {
                    allSame = false;
                    std::printf("   !! %s %s: output differs from %s\n", variants[vi].name.c_str(), m.c_str(), variants[0].name.c_str());
                }
            }
            std::vector<std::vector<double>> gbs(variants.size());
            for (uint32_t r = 0; r < rounds; ++r)
                for (size_t k = 0; k < variants.size(); ++k) {
                    const size_t vi = (k + r) % variants.size();   // rotate the order: power capping favours whoever runs first
                    for (uint32_t w = 0; w < 10; ++w) launch(variants[vi], w);
                    CK(cudaStreamSynchronize(stream));
                    CK(cudaEventRecord(e0, stream));
                    for (uint32_t it = 0; it < iters; ++it) launch(variants[vi], it);
                    CK(cudaEventRecord(e1, stream));
                    CK(cudaStreamSynchronize(stream));
                    float ms = 0;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    gbs[vi].push_back(px * bpp * iters / (ms * 1e-3) / 1e9);
                }
            js << (firstMode ? "" : ", ") << "\"" << m << "\": {";
            firstMode = false;
            for (size_t vi = 0; vi < variants.size(); ++vi) {
                double best = 0, mean = 0;
                for (double g : gbs[vi]) { best = std::max(best, g); mean += g / gbs[vi].size(); }
                std::vector<double> sorted = gbs[vi];
                std::sort(sorted.begin(), sorted.end());
                const double median = sorted[sorted.size() / 2];
                std::printf("   %-5s %-28s median %7.0f  mean %7.0f  best %7.0f GB/s  (%.2f us per launch)\n", m.c_str(), variants[vi].name.c_str(),
                            median, mean, best, px * bpp / (median * 1e9) * 1e6);
                js << (vi ? ", " : "") << "\"" << variants[vi].name << "\": {\"median_gbs\": " << median << ", \"mean_gbs\": " << mean
                   << ", \"best_gbs\": " << best << "}";
            }
            js << "}";
        }
        js << "}";
    }
    js << ", \"all_variants_same_bytes\": " << (allSame ? "true" : "false") << ", \"iters\": " << iters << ", \"rounds\": " << rounds << "}";
    if (!jsonPath.empty()) std::ofstream(jsonPath) << js.str() << "\n";
    std::printf("all variants same bytes: %s\n", allSame ? "yes" : "NO");
    return allSame ? 0 : 5;
}
