// membench.cu -- what HBM bandwidth can a streaming kernel with the encoders' access pattern reach?
// Diagnostic tool (not part of libgoofy_b200.so).  Read-mostly probes with the same 8:1 read:write
// ratio as the encoders, so the roofline fraction in bench.py can be read against a like-for-like
// ceiling rather than against a 1:1 copy.
//   linear   : each thread reads 64 contiguous bytes (4 x LDG.128), writes 8
//   rows4    : each thread reads 16 bytes from 4 rows `stride` apart (the encoders' pattern), writes 8
//   copy     : 16 B in, 16 B out (STREAM copy, what MEASURED_PEAKS.json's hbm_gbs is)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint4 ldg_na(const void* p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(256) k_rows4(const uint8_t* src, uint8_t* dst, uint32_t bw, uint32_t stride)
{
    const uint32_t bx = blockIdx.x * 256 + threadIdx.x, by = blockIdx.y;
    const uint8_t* s = src + (size_t)by * 4 * stride + (size_t)bx * 16;
    uint4 a = ldg_na(s), b = ldg_na(s + stride), c = ldg_na(s + 2 * (size_t)stride), d = ldg_na(s + 3 * (size_t)stride);
    uint2 o;
    o.x = a.x ^ b.y ^ c.z ^ d.w ^ a.z ^ b.w ^ c.x ^ d.y;
    o.y = a.y ^ b.z ^ c.w ^ d.x ^ a.w ^ b.x ^ c.y ^ d.z;
    *(uint2*)(dst + ((size_t)by * bw + bx) * 8) = o;
}

// variants of rows4 to look for headroom above the plain pattern
__device__ __forceinline__ uint64_t evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// 256-bit load (sm_100): 32 bytes per thread per instruction
struct u32x8 { uint32_t v[8]; };
__device__ __forceinline__ u32x8 ldg256(const void* p)
{
    u32x8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
}

// two horizontally adjacent blocks per thread, 4 x 256-bit loads, one 16-byte store
__global__ void __launch_bounds__(128) k_rows4x256(const uint8_t* src, uint8_t* dst, uint32_t bw, uint32_t stride)
{
    const uint32_t bx2 = blockIdx.x * 128 + threadIdx.x, by = blockIdx.y;
    const uint8_t* s = src + (size_t)by * 4 * stride + (size_t)bx2 * 32;
    u32x8 a = ldg256(s), b = ldg256(s + stride), c = ldg256(s + 2 * (size_t)stride), d = ldg256(s + 3 * (size_t)stride);
    uint4 o;
    o.x = a.v[0] ^ b.v[1] ^ c.v[2] ^ d.v[3]; o.y = a.v[1] ^ b.v[2] ^ c.v[3] ^ d.v[0];
    o.z = a.v[4] ^ b.v[5] ^ c.v[6] ^ d.v[7]; o.w = a.v[5] ^ b.v[6] ^ c.v[7] ^ d.v[4];
    *(uint4*)(dst + ((size_t)by * bw + bx2 * 2) * 8) = o;
}

template <int HINT>
__device__ __forceinline__ uint4 ldg_hint(const void* p)
{
    uint4 v;
    if (HINT == 0) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (HINT == 1) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (HINT == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (HINT == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(evict_first_policy()));
    if (HINT == 4) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int HINT, int TPB>
__global__ void __launch_bounds__(TPB) k_rows4v(const uint8_t* src, uint8_t* dst, uint32_t bw, uint32_t stride)
{
    const uint32_t bx = blockIdx.x * TPB + threadIdx.x, by = blockIdx.y;
    const uint8_t* s = src + (size_t)by * 4 * stride + (size_t)bx * 16;
    uint4 a = ldg_hint<HINT>(s), b = ldg_hint<HINT>(s + stride), c = ldg_hint<HINT>(s + 2 * (size_t)stride), d = ldg_hint<HINT>(s + 3 * (size_t)stride);
    uint2 o;
    o.x = a.x ^ b.y ^ c.z ^ d.w ^ a.z ^ b.w ^ c.x ^ d.y;
    o.y = a.y ^ b.z ^ c.w ^ d.x ^ a.w ^ b.x ^ c.y ^ d.z;
    *(uint2*)(dst + ((size_t)by * bw + bx) * 8) = o;
}

// the dual-output kernel's pattern: 64 B read, two 8-byte blocks written to two output streams (4:1 read:write)
__global__ void __launch_bounds__(256) k_rows4dual(const uint8_t* src, uint8_t* dst, uint8_t* dst2, uint32_t bw, uint32_t stride)
{
    const uint32_t bx = blockIdx.x * 256 + threadIdx.x, by = blockIdx.y;
    const uint8_t* s = src + (size_t)by * 4 * stride + (size_t)bx * 16;
    uint4 a = ldg_hint<0>(s), b = ldg_hint<0>(s + stride), c = ldg_hint<0>(s + 2 * (size_t)stride), d = ldg_hint<0>(s + 3 * (size_t)stride);
    uint2 o, o2;
    o.x = a.x ^ b.y ^ c.z ^ d.w; o.y = a.z ^ b.w ^ c.x ^ d.y;
    o2.x = a.y ^ b.z ^ c.w ^ d.x; o2.y = a.w ^ b.x ^ c.y ^ d.z;
    *(uint2*)(dst + ((size_t)by * bw + bx) * 8) = o;
    *(uint2*)(dst2 + ((size_t)by * bw + bx) * 8) = o2;
}

// the decoder's pattern, the encoder's mirrored: 8 bytes read, four 16-byte row stores (1:8 read:write)
__global__ void __launch_bounds__(256) k_rows4decode(const uint8_t* blocks, uint8_t* rgba, uint32_t bw, uint32_t stride)
{
    const uint32_t bx = blockIdx.x * 256 + threadIdx.x, by = blockIdx.y;
    const uint2 b = *(const uint2*)(blocks + ((size_t)by * bw + bx) * 8);
    uint8_t* o = rgba + (size_t)by * 4 * stride + (size_t)bx * 16;
    const uint4 v0 = make_uint4(b.x, b.y, b.x ^ b.y, ~b.x), v1 = make_uint4(b.y, b.x, ~b.y, b.x + b.y);
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(o), "r"(v0.x), "r"(v0.y), "r"(v0.z), "r"(v0.w) : "memory");
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(o + stride), "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w) : "memory");
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(o + 2 * (size_t)stride), "r"(v0.y), "r"(v0.x), "r"(v0.w), "r"(v0.z) : "memory");
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(o + 3 * (size_t)stride), "r"(v1.y), "r"(v1.x), "r"(v1.w), "r"(v1.z) : "memory");
}

// two vertically adjacent blocks per thread: 8 row loads in flight
__global__ void __launch_bounds__(256) k_rows8(const uint8_t* src, uint8_t* dst, uint32_t bw, uint32_t stride)
{
    const uint32_t bx = blockIdx.x * 256 + threadIdx.x, by = blockIdx.y * 2;
    const uint8_t* s = src + (size_t)by * 4 * stride + (size_t)bx * 16;
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ldg_na(s + (size_t)i * stride);
    uint2 o, o2;
    o.x = v[0].x ^ v[1].y ^ v[2].z ^ v[3].w; o.y = v[0].y ^ v[1].z ^ v[2].w ^ v[3].x;
    o2.x = v[4].x ^ v[5].y ^ v[6].z ^ v[7].w; o2.y = v[4].y ^ v[5].z ^ v[6].w ^ v[7].x;
    *(uint2*)(dst + ((size_t)by * bw + bx) * 8) = o;
    *(uint2*)(dst + ((size_t)(by + 1) * bw + bx) * 8) = o2;
}

// read only (no output stream at all): the pure-read ceiling
__global__ void __launch_bounds__(256) k_readonly(const uint8_t* src, uint32_t* sink, uint32_t stride)
{
    const uint32_t bx = blockIdx.x * 256 + threadIdx.x, by = blockIdx.y;
    const uint8_t* s = src + (size_t)by * 4 * stride + (size_t)bx * 16;
    uint4 a = ldg_na(s), b = ldg_na(s + stride), c = ldg_na(s + 2 * (size_t)stride), d = ldg_na(s + 3 * (size_t)stride);
    uint32_t x = a.x ^ b.y ^ c.z ^ d.w ^ a.z ^ b.w ^ c.x ^ d.y ^ a.y ^ b.z ^ c.w ^ d.x ^ a.w ^ b.x ^ c.y ^ d.z;
    if (x == 0x12345678u) sink[0] = x;
}

__global__ void __launch_bounds__(256) k_linear(const uint8_t* src, uint8_t* dst)
{
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t warp = t >> 5, lane = t & 31;
    const uint8_t* s = src + warp * 2048 + lane * 16;  // a warp reads 2 KiB contiguous as 4 x 512 B
    uint4 a = ldg_na(s), b = ldg_na(s + 512), c = ldg_na(s + 1024), d = ldg_na(s + 1536);
    uint2 o;
    o.x = a.x ^ b.y ^ c.z ^ d.w ^ a.z ^ b.w ^ c.x ^ d.y;
    o.y = a.y ^ b.z ^ c.w ^ d.x ^ a.w ^ b.x ^ c.y ^ d.z;
    *(uint2*)(dst + t * 8) = o;
}

__global__ void __launch_bounds__(256) k_copy(const uint4* src, uint4* dst, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) dst[i] = src[i];
}

int main(int argc, char** argv)
{
    // default: one 8192x8192 texture per launch (302 MB); "membench 4" = four textures per launch (1.2 GB, bench.py's step)
    const uint32_t tall = argc > 1 ? (uint32_t)atoi(argv[1]) : 1u;
    const uint32_t W = 8192, H = 8192 * tall, bw = W / 4, bh = H / 4, stride = W * 4;
    const size_t inBytes = (size_t)W * H * 4, outBytes = inBytes / 8;
    const int NBUF = tall > 1 ? 2 : 4;
    uint8_t* src[NBUF]; uint8_t* dst[NBUF];
    for (int i = 0; i < NBUF; ++i) { cudaMalloc(&src[i], inBytes); cudaMalloc(&dst[i], inBytes); cudaMemset(src[i], i + 1, inBytes); }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, double bytes, auto launch) {
        for (int i = 0; i < 8; ++i) launch(i % NBUF);
        cudaDeviceSynchronize();
        const int iters = 100;
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) launch(i % NBUF);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-11s %8.2f us/launch  %8.1f GB/s\n", name, ms / iters * 1e3, bytes / (ms / iters * 1e-3) / 1e9);
    };
    timeit("rows4", (double)inBytes + outBytes, [&](int b) { k_rows4<<<dim3(bw / 256, bh), 256>>>(src[b], dst[b], bw, stride); });
    timeit("rows4/nc", (double)inBytes + outBytes, [&](int b) { k_rows4v<0, 256><<<dim3(bw / 256, bh), 256>>>(src[b], dst[b], bw, stride); });
    timeit("rows4/256B", (double)inBytes + outBytes, [&](int b) { k_rows4v<1, 256><<<dim3(bw / 256, bh), 256>>>(src[b], dst[b], bw, stride); });
    timeit("rows4/128B", (double)inBytes + outBytes, [&](int b) { k_rows4v<2, 256><<<dim3(bw / 256, bh), 256>>>(src[b], dst[b], bw, stride); });
    timeit("rows4/evf", (double)inBytes + outBytes, [&](int b) { k_rows4v<3, 256><<<dim3(bw / 256, bh), 256>>>(src[b], dst[b], bw, stride); });
    timeit("rows4/cs", (double)inBytes + outBytes, [&](int b) { k_rows4v<4, 256><<<dim3(bw / 256, bh), 256>>>(src[b], dst[b], bw, stride); });
    timeit("rows4/t128", (double)inBytes + outBytes, [&](int b) { k_rows4v<0, 128><<<dim3(bw / 128, bh), 128>>>(src[b], dst[b], bw, stride); });
    timeit("rows4/t512", (double)inBytes + outBytes, [&](int b) { k_rows4v<0, 512><<<dim3(bw / 512, bh), 512>>>(src[b], dst[b], bw, stride); });
    timeit("rows4/dual", (double)inBytes + 2.0 * outBytes, [&](int b) { k_rows4dual<<<dim3(bw / 256, bh), 256>>>(src[b], dst[b], dst[b] + outBytes, bw, stride); });
    timeit("rows4/decode", (double)inBytes + outBytes, [&](int b) { k_rows4decode<<<dim3(bw / 256, bh), 256>>>(dst[b], src[b], bw, stride); });
    for (int i = 0; i < NBUF; ++i) cudaMemset(src[i], i + 1, inBytes);
    timeit("rows4x256", (double)inBytes + outBytes, [&](int b) { k_rows4x256<<<dim3(bw / 256, bh), 128>>>(src[b], dst[b], bw, stride); });
    timeit("rows8", (double)inBytes + outBytes, [&](int b) { k_rows8<<<dim3(bw / 256, bh / 2), 256>>>(src[b], dst[b], bw, stride); });
    timeit("readonly", (double)inBytes, [&](int b) { k_readonly<<<dim3(bw / 256, bh), 256>>>(src[b], (uint32_t*)dst[b], stride); });
    timeit("linear", (double)inBytes + outBytes, [&](int b) { k_linear<<<(unsigned)(inBytes / 64 / 256), 256>>>(src[b], dst[b]); });
    timeit("copy", 2.0 * inBytes, [&](int b) { k_copy<<<148 * 16, 256>>>((const uint4*)src[b], (uint4*)dst[b], inBytes / 16); });
    {
        float best = 1e9f;
        for (int i = 0; i < 10; ++i) {
            cudaEventRecord(e0); cudaMemcpyAsync(dst[i % NBUF], src[i % NBUF], inBytes, cudaMemcpyDeviceToDevice); cudaEventRecord(e1);
            cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-8s %8.2f us/launch  %8.1f GB/s (cudaMemcpy D2D, best of 10)\n", "memcpy", best * 1e3, 2.0 * inBytes / (best * 1e-3) / 1e9);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
