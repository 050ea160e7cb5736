// encode_kernels.cuh -- the global kernels around block_codec.cuh: image load/tiling layer
// and output block writer.
//
// Replaces the row/tile loops of goofy::compressDXT1/ETC1 (GoofyTC/goofy_tc.h:1514-1524,
// :1545-1555) and the tile fetch / dword stores of goofySimdEncode (:1077-1099, :1332-1356,
// :1476-1492).  One thread owns one 4x4 block:
//   load   4 x LDG.128 (one per block row).  A warp covers 32 adjacent blocks, so each of the
//          four requests is 512 contiguous bytes -- four full 128-byte lines, fully coalesced.
//   store  one 8-byte block per thread -> 256 contiguous bytes per warp.
// Inputs are streamed once, so loads bypass L1 allocation and stores are streaming.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "block_codec.cuh"

namespace gb {

enum Mode : int { kDxt1 = 0, kEtc1 = 1, kDual = 2 };

#ifndef GB_TPB
#define GB_TPB 256     // threads per CTA of the direct / row-walking kernels (x a power of two, x*y == GB_TPB)
#endif
// Resident 256-thread CTAs per SM each flavour is compiled for (8 -> 32 registers, 6 -> 40, 5 -> 48).
#ifndef GB_CTAS_DXT1
#define GB_CTAS_DXT1 8
#endif
#ifndef GB_CTAS_ETC1
#define GB_CTAS_ETC1 6
#endif
#ifndef GB_CTAS_DUAL
#define GB_CTAS_DUAL 6
#endif
// Resident CTAs per SM of the cp.async ring kernels (33 KB of shared memory each)
#ifndef GB_ASYNC_CTAS_DUAL
#define GB_ASYNC_CTAS_DUAL 5
#endif
#ifndef GB_ASYNC_CTAS_SINGLE
#define GB_ASYNC_CTAS_SINGLE 5
#endif
#define GB_ASYNC_CTAS(mode) ((mode) == 2 ? GB_ASYNC_CTAS_DUAL : GB_ASYNC_CTAS_SINGLE)
// Resident CTAs per SM of the row-walking ETC1s relaxed-shape kernel (5 -> 48 registers, no spills; 6 -> 40 registers, 24 bytes spilled)
#ifndef GB_RELAXED_ETC1_CTAS
#define GB_RELAXED_ETC1_CTAS 5
#endif
// Selector-gathering scheme of the DXT1 kernel (block_codec.cuh `Selectors`); -D overridable for A/B runs.
// The ETC1s and dual-output kernels always use the flag-byte scheme.
#ifndef GB_SEL_DXT1
#define GB_SEL_DXT1 kSelLanes
#endif

constexpr int ctas_per_sm(int mode) { return mode == 0 ? GB_CTAS_DXT1 : mode == 1 ? GB_CTAS_ETC1 : GB_CTAS_DUAL; }
// The packed-RGB kernels hold twelve loaded words next to the sixteen pixel words they widen them to: the DXT1 flavour
// spills 20-40 bytes at 32 registers, so it runs 6 CTAs per SM (40 registers) like the other two.
#ifndef GB_RGB24_CTAS_DXT1
#define GB_RGB24_CTAS_DXT1 6
#endif
#ifndef GB_RGB24_CTAS_ETC1
#define GB_RGB24_CTAS_ETC1 GB_CTAS_ETC1
#endif
#ifndef GB_RGB24_CTAS_DUAL
#define GB_RGB24_CTAS_DUAL 5   // 5418 -> 5615 GB/s against 6 (session V); DXT1 at 5 / 7 / 8: 6401 / 6096 / 6096 against 6815 at 6
#endif
constexpr int rgb24_ctas_per_sm(int mode) { return mode == 0 ? GB_RGB24_CTAS_DXT1 : mode == 1 ? GB_RGB24_CTAS_ETC1 : GB_RGB24_CTAS_DUAL; }

struct EncodeParams {
    const uint8_t* src;
    uint8_t* dst;        // DXT1 (kDxt1, kDual) or ETC1 (kEtc1) blocks
    uint8_t* dst2;       // ETC1 blocks for kDual
    uint32_t bw, bh;     // image size in blocks
    uint32_t stride;     // bytes between pixel rows
    uint32_t by0;        // first block row handled by this launch (grid.y chunking)
    uint32_t firstWave;  // CTAs (in launch order) that warm L2 with their first tile before griddepcontrol.wait; 0 = none
    uint32_t prefetchNext;  // row-walking kernel: warm L2 with the thread's NEXT block row while encoding this one (short launches)
    uint64_t srcPitch;   // bytes between images
    uint64_t dstPitch;
};

// ETC1 control words by clamped brightness range, staged in shared memory by every CTA from a table in device
// memory.  The table is a `__device__ const` object with a CONSTANT initialiser (built at compile time from the
// seven thresholds), so it is part of the module image: there is no per-device initialisation kernel, the first
// call on a device neither synchronises it nor breaks a stream capture.  Same content as the reference's table
// (goofy_tc.h:1040-1057), generated, not copied.  [1] = the float-reference flavour's thresholds
// (Src/goofy_tc_reference.cpp:684-720).  (Computing the entries in every CTA instead was measured 3-5 % slower on
// the ETC1s / dual-output kernels: seventeen more ALU-pipe instructions per thread and CTA.)
struct ControlTables {
    uint32_t word[2][256];
};
constexpr ControlTables make_control_tables()
{
    ControlTables t = {};
    for (uint32_t r = 0; r < 256u; ++r) {
        const uint32_t cw = (r >= 22u) + (r >= 44u) + (r >= 74u) + (r >= 106u) + (r >= 152u) + (r >= 182u) + (r >= 254u);
        const uint32_t cwRef = (r > 10u) + (r > 21u) + (r > 36u) + (r > 52u) + (r > 75u) + (r > 90u) + (r > 126u);
        t.word[0][r] = (cw * 36u + 3u) << 24;
        t.word[1][r] = (cwRef * 36u + 3u) << 24;
    }
    return t;
}
__device__ const ControlTables g_controlTables = make_control_tables();

template <bool REF, int NTHREADS>
__device__ __forceinline__ void stage_control_lut(uint32_t* lut, uint32_t t)
{
#pragma unroll
    for (uint32_t i = t; i < 256u; i += NTHREADS) lut[i] = g_controlTables.word[REF ? 1 : 0][i];
}

// Programmatic dependent launch: when the host launches with programmatic stream serialisation, the
// CTAs of this kernel may become resident while the previous kernel on the stream is still draining.
// `pdl_wait` blocks until that kernel has fully completed and its writes are visible, so ordering is
// exactly what plain stream order gives -- only the launch latency and the CTA ramp-up are hidden
// (back-to-back 45 us launches otherwise lose ~7 % to it).  Both are no-ops for a normal launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Warm L2 with one 128-byte line.  Issued by the first wave of CTAs BEFORE griddepcontrol.wait: a prefetch has no
// architectural effect (L2 is the coherence point, so a line the previous kernel writes afterwards is still read
// correctly), but DRAM starts streaming this launch's first tiles while the previous kernel is still draining its
// last wave, instead of idling until every CTA of it has retired.
__device__ __forceinline__ void prefetch_l2(const uint8_t* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ uint4 load_row(const uint8_t* p)
{
    uint4 v;
#ifndef GB_LOAD_PTX
#define GB_LOAD_PTX "ld.global.nc.v4.u32"
#endif
    asm volatile(GB_LOAD_PTX " {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ void store_block(uint8_t* p, uint32_t w0, uint32_t w1)
{
#ifndef GB_STORE_PTX
#define GB_STORE_PTX "st.global.cs.v2.u32"
#endif
    asm volatile(GB_STORE_PTX " [%0], {%1,%2};" ::"l"(p), "r"(w0), "r"(w1) : "memory");
}

// Encode the 16 pixels of one block in MODE and write the 8-byte result(s).
template <int MODE>
__device__ __forceinline__ void encode_and_store(const uint4& r0, const uint4& r1, const uint4& r2, const uint4& r3,
                                                 const uint32_t* lut, uint8_t* dst, uint8_t* dst2)
{
    const uint32_t p[16] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w,
                            r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
    const BlockFront f = analyse(p);
    uint32_t w0, w1;
    if (MODE == kDxt1) {
        encode_dxt1<GB_SEL_DXT1>(p, f, w0, w1);
        store_block(dst, w0, w1);
    } else if (MODE == kEtc1) {
        encode_etc1(p, f, lut, w0, w1);
        store_block(dst, w0, w1);
    } else {
        uint32_t e0, e1;
        encode_both(p, f, lut, w0, w1, e0, e1);
        store_block(dst, w0, w1);
        store_block(dst2, e0, e1);
    }
}

// The same for either flavour: FLAVOUR 0 = SSE2-exact (goofy::), 1 = float-reference-exact (goofyRef::, single codecs only).
template <int MODE, int FLAVOUR>
__device__ __forceinline__ void encode_and_store_flavour(const uint4& r0, const uint4& r1, const uint4& r2, const uint4& r3,
                                                         const uint32_t* lut, uint8_t* dst, uint8_t* dst2)
{
    if (FLAVOUR == 0) {
        encode_and_store<MODE>(r0, r1, r2, r3, lut, dst, dst2);
    } else {
        const uint32_t p[16] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w,
                                r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
        // minimum brightness range: 8 for DXT1 (Src/goofy_tc_reference.cpp:667), 16 for ETC1S (:766), times 4
        const RefFront f = analyse_ref(p, MODE == kDxt1 ? 32u : 64u);
        uint32_t w0, w1;
        if (MODE == kDxt1) encode_dxt1_ref(p, f, w0, w1);
        else encode_etc1_ref(p, f, lut, w0, w1);
        store_block(dst, w0, w1);
    }
}

// Resident CTAs per SM: 8 (32 registers) for DXT1, 6 (40 registers) for ETC1s and dual-output (ctas_per_sm).
// WIDE = false: every byte offset inside one image fits 32 bits (the launcher checks), which
// keeps the address arithmetic to a handful of 32-bit ops; WIDE = true is the same kernel with
// 64-bit offsets for images of 4 GiB and more.
// PITCHED = true adds blockIdx.z * pitch for batches whose images are not back to back (batches
// that ARE back to back are launched as one tall image, so the common case pays nothing).
// SHORT = the instantiation for short launches: its first wave of CTAs warms L2 before griddepcontrol.wait (prefetch_l2).
// Long launches run the instantiation without that code: two dozen instructions per block are worth 1-5 % to the
// kernels that sit near their issue limit when the box is power-capped.
template <int MODE, bool WIDE, bool PITCHED, bool SHORT>
__global__ void __launch_bounds__(GB_TPB, ctas_per_sm(MODE) * 256 / GB_TPB) encode_direct_kernel(const EncodeParams P)
{
    __shared__ uint32_t lut[MODE == kDxt1 ? 1 : 256];
    pdl_launch_dependents();
    // the launcher always uses GB_TPB threads (x a power of two, x*y == GB_TPB)
    if (MODE != kDxt1) stage_control_lut<false, GB_TPB>(lut, threadIdx.y * blockDim.x + threadIdx.x);

    const uint32_t bx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t by = P.by0 + blockIdx.y * blockDim.y + threadIdx.y;
    const bool live = bx < P.bw && by < P.bh;

    typedef typename std::conditional<WIDE, uint64_t, uint32_t>::type off_t;
    const off_t o0 = (off_t)by * ((off_t)4u * P.stride) + (off_t)(bx * 16u);
    const off_t o1 = o0 + P.stride, o2 = o1 + P.stride, o3 = o2 + P.stride;
    const uint8_t* src = P.src;
    uint8_t* dst = P.dst;
    uint8_t* dst2 = P.dst2;
    if (PITCHED) {
        src += (uint64_t)blockIdx.z * P.srcPitch;
        dst += (uint64_t)blockIdx.z * P.dstPitch;
        if (MODE == kDual) dst2 += (uint64_t)blockIdx.z * P.dstPitch;
    }
    if (SHORT && P.firstWave != 0u) {
        const uint32_t cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        if (cta < P.firstWave && live && (threadIdx.x & 7u) == 0u) {   // eight neighbouring threads share a 128-byte line
            prefetch_l2(src + o0);
            prefetch_l2(src + o1);
            prefetch_l2(src + o2);
            prefetch_l2(src + o3);
        }
    }
    pdl_wait();
    if (MODE != kDxt1) __syncthreads();
    if (!live) return;
    const uint4 r0 = load_row(src + o0);
    const uint4 r1 = load_row(src + o1);
    const uint4 r2 = load_row(src + o2);
    const uint4 r3 = load_row(src + o3);
    const off_t o = ((off_t)by * P.bw + bx) * 8u;
    encode_and_store<MODE>(r0, r1, r2, r3, lut, dst + o, MODE == kDual ? dst2 + o : nullptr);
}

// Row-walking variant: blockIdx.x picks the column strip and every CTA walks down its image in steps of
// gridDim.y * blockDim.y block rows (the launcher sizes the grid to a few rows per CTA, never fewer CTAs than
// are resident at once).  PITCHED: blockIdx.z picks the image of a batch laid out at fixed pitches.  The per-thread set-up
// (indices, control table, constants) is paid once and each further block costs only the
// pointer bumps -- about 25 fewer instructions per block than one-shot CTAs.
template <int MODE, bool WIDE, bool PITCHED, bool SHORT>
__global__ void __launch_bounds__(GB_TPB, ctas_per_sm(MODE) * 256 / GB_TPB) encode_rows_kernel(const EncodeParams P)
{
    __shared__ uint32_t lut[MODE == kDxt1 ? 1 : 256];
    pdl_launch_dependents();
    if (MODE != kDxt1) stage_control_lut<false, GB_TPB>(lut, threadIdx.y * blockDim.x + threadIdx.x);
    const uint8_t* src = P.src;
    uint8_t* dst = P.dst;
    uint8_t* dst2 = P.dst2;
    if (PITCHED) {   // grid.z = image of a batch laid out at fixed pitches
        src += (uint64_t)blockIdx.z * P.srcPitch;
        dst += (uint64_t)blockIdx.z * P.dstPitch;
        if (MODE == kDual) dst2 += (uint64_t)blockIdx.z * P.dstPitch;
    }
    const uint32_t bx = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t by = blockIdx.y * blockDim.y + threadIdx.y;
    const bool live = bx < P.bw && by < P.bh;
    typedef typename std::conditional<WIDE, uint64_t, uint32_t>::type off_t;
    if (SHORT && P.firstWave != 0u) {
        const uint32_t cta = blockIdx.x + gridDim.x * (blockIdx.y + (PITCHED ? gridDim.y * blockIdx.z : 0u));
        if (cta < P.firstWave && live && (threadIdx.x & 7u) == 0u) {
            const off_t o0 = (off_t)by * ((off_t)4u * P.stride) + (off_t)(bx * 16u);
            prefetch_l2(src + o0);
            prefetch_l2(src + o0 + P.stride);
            prefetch_l2(src + o0 + 2u * (off_t)P.stride);
            prefetch_l2(src + o0 + 3u * (off_t)P.stride);
        }
    }
    pdl_wait();
    if (MODE != kDxt1) __syncthreads();
    if (!live) return;

    const uint32_t rowStep = gridDim.y * blockDim.y;
    // Only `by` is carried round the loop; offsets are rebuilt from it (a few IMADs on the
    // otherwise idle multiply pipe) so the loop state stays inside the 32-register budget.
#pragma unroll 1
    for (; by < P.bh; by += rowStep) {
        const off_t o0 = (off_t)by * ((off_t)4u * P.stride) + (off_t)(bx * 16u);
        const off_t o1 = o0 + P.stride, o2 = o1 + P.stride, o3 = o2 + P.stride;
        if (SHORT && P.prefetchNext != 0u && (threadIdx.x & 7u) == 0u && by + rowStep < P.bh) {
            // a short launch never reaches the steady state in which the warps of an SM sit in different phases: its
            // CTAs load together and encode together, so DRAM idles while they encode unless somebody keeps it busy
            const off_t n0 = o0 + (off_t)rowStep * ((off_t)4u * P.stride);
            prefetch_l2(src + n0);
            prefetch_l2(src + n0 + P.stride);
            prefetch_l2(src + n0 + 2u * (off_t)P.stride);
            prefetch_l2(src + n0 + 3u * (off_t)P.stride);
        }
        const uint4 r0 = load_row(src + o0);
        const uint4 r1 = load_row(src + o1);
        const uint4 r2 = load_row(src + o2);
        const uint4 r3 = load_row(src + o3);
        const off_t o = ((off_t)by * P.bw + bx) * 8u;
        encode_and_store<MODE>(r0, r1, r2, r3, lut, dst + o, MODE == kDual ? dst2 + o : nullptr);
    }
}

// Row-walking CTAs with an asynchronous double buffer: while a thread encodes its block of the current
// row group, the four 16-byte rows of its NEXT block are already on their way into shared memory
// (cp.async / LDGSTS, no registers held).  Every thread only ever reads what its own cp.async wrote, so
// the ring needs no barrier -- just cp.async.wait_group.  This removes most of the load-wait stalls
// that ncu shows on the first use of the loaded pixels (41 % of stall samples in the plain kernel).
__device__ __forceinline__ void cp_async16(uint32_t smemAddr, const uint8_t* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smemAddr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int MODE, bool WIDE>
__global__ void __launch_bounds__(GB_TPB, GB_ASYNC_CTAS(MODE)) encode_rows_async_kernel(const EncodeParams P)
{
    __shared__ uint32_t lut[MODE == kDxt1 ? 1 : 256];
    __shared__ __align__(16) uint4 ring[2][4][GB_TPB];  // [stage][pixel row][thread]
    pdl_launch_dependents();
    const uint32_t t = threadIdx.y * blockDim.x + threadIdx.x;
    if (MODE != kDxt1) {
        stage_control_lut<false, GB_TPB>(lut, t);
        __syncthreads();
    }
    pdl_wait();
    const uint32_t bx = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t by = blockIdx.y * blockDim.y + threadIdx.y;
    if (bx >= P.bw || by >= P.bh) return;

    typedef typename std::conditional<WIDE, uint64_t, uint32_t>::type off_t;
    const uint32_t rowStep = gridDim.y * blockDim.y;
    const uint32_t ringBase = (uint32_t)__cvta_generic_to_shared(&ring[0][0][t]);
    auto prefetch = [&](uint32_t row, uint32_t stage) {
        const off_t o0 = (off_t)row * ((off_t)4u * P.stride) + (off_t)(bx * 16u);
        const off_t o1 = o0 + P.stride, o2 = o1 + P.stride, o3 = o2 + P.stride;
        const uint32_t s = ringBase + stage * (uint32_t)sizeof(ring[0]);
        cp_async16(s, P.src + o0);
        cp_async16(s + (uint32_t)sizeof(ring[0][0]), P.src + o1);
        cp_async16(s + 2u * (uint32_t)sizeof(ring[0][0]), P.src + o2);
        cp_async16(s + 3u * (uint32_t)sizeof(ring[0][0]), P.src + o3);
    };
    prefetch(by, 0u);
    cp_async_commit();
    uint32_t stage = 0;
#pragma unroll 1
    for (;;) {
        const uint32_t byNext = by + rowStep;
        if (byNext < P.bh) prefetch(byNext, stage ^ 1u);
        cp_async_commit();          // (possibly empty) group: keeps "all but the newest" == the current stage
        cp_async_wait<1>();
        const uint4 r0 = ring[stage][0][t], r1 = ring[stage][1][t], r2 = ring[stage][2][t], r3 = ring[stage][3][t];
        const off_t o = ((off_t)by * P.bw + bx) * 8u;
        encode_and_store<MODE>(r0, r1, r2, r3, lut, P.dst + o, MODE == kDual ? P.dst2 + o : nullptr);
        if (byNext >= P.bh) break;
        by = byNext;
        stage ^= 1u;
    }
}

// Float-reference flavour (goofyRef::, block_codec.cuh "float-reference flavour"), any width that is a multiple of 4;
// the control table is goofyRef's, by brightRange.  grid.z = image of a batch at fixed pitches.
// ROWS = false: one-shot CTAs (what the DXT1 flavour runs: it is HBM-bound either way, like its SSE2-exact twin).
// ROWS = true: every CTA walks down its image in steps of gridDim.y * blockDim.y block rows, so the control-table
// staging and the index set-up are paid once per few blocks (the ETC1s flavour: 6.06 -> TB/s as one-shot CTAs, r01f).
template <int CODEC, bool ROWS>
__global__ void __launch_bounds__(256, CODEC == kDxt1 ? 8 : 6) encode_floatref_kernel(const EncodeParams P)
{
    __shared__ uint32_t lut[CODEC == kDxt1 ? 1 : 256];
    pdl_launch_dependents();
    if (CODEC != kDxt1) stage_control_lut<true, 256>(lut, threadIdx.y * blockDim.x + threadIdx.x);
    pdl_wait();
    if (CODEC != kDxt1) __syncthreads();
    const uint32_t bx = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t by = P.by0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (bx >= P.bw || by >= P.bh) return;
    const uint8_t* src = P.src + (uint64_t)blockIdx.z * P.srcPitch + (uint64_t)bx * 16u;
    uint8_t* dst = P.dst + (uint64_t)blockIdx.z * P.dstPitch + (uint64_t)bx * 8u;
#pragma unroll 1
    do {
        const uint8_t* s = src + (uint64_t)by * 4u * P.stride;
        const uint4 r0 = load_row(s);
        const uint4 r1 = load_row(s + P.stride);
        const uint4 r2 = load_row(s + 2ull * P.stride);
        const uint4 r3 = load_row(s + 3ull * P.stride);
        encode_and_store_flavour<CODEC, 1>(r0, r1, r2, r3, lut, dst + (uint64_t)by * P.bw * 8u, nullptr);
    } while (ROWS && (by += gridDim.y * blockDim.y) < P.bh);
}

// Packed RGB8 input: 3 bytes per pixel, no alpha byte (the encoders ignore alpha, GoofyTC/goofy_tc.h:297, so a
// caller that holds RGB data -- a decoded PNG/JPEG, or the host path's alpha-dropping staging copy, rgb_pack.h --
// need not widen it to RGBA first: 3.5 instead of 4.5 bytes of traffic per pixel).  A block row of one block is 12
// bytes at a 4-byte-aligned address: three 32-bit loads (a warp covers 384 contiguous bytes per pixel row with them;
// L1 merges the three requests per line), widened to the four pixel words of the RGBA kernels with two byte permutes
// and a shift.  The fourth byte of every pixel word is whatever followed it, which the arithmetic never looks at
// (tests/test_gpu_parity.py::test_alpha_is_ignored).  Row-walking CTAs, grid.z = image of a batch at fixed pitches.
// (A warp-cooperative variant -- three coalesced 128-bit loads per lane into a warp-private shared-memory tile, read
// back as twelve conflict-free LDS.32 -- was measured 16-19 % SLOWER on device memory, 5661 vs 6713 GB/s DXT1, and no
// different on pinned host memory read over PCIe: the extra STS / LDS / __syncwarp instructions cost more than the
// narrower requests; session O, profiles/r02_rgb24_sessions.md.)
__device__ __forceinline__ uint32_t load_word(const uint8_t* p)
{
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// w0 = R0 G0 B0 R1 | w1 = G1 B1 R2 G2 | w2 = B2 R3 G3 B3  ->  four pixel words (byte 3 of each: don't care)
// (The same widening as IMAD.HI + IMAD on the multiply pipe -- funnel shifts are multiplications by run-time powers of
// two -- takes the twelve instructions off the integer ALU pipe, which ncu shows 74-78 % busy in these kernels, and
// measured 2-3 % SLOWER: session W.)
__device__ __forceinline__ uint4 load_row_rgb24(const uint8_t* p)
{
    const uint32_t w0 = load_word(p), w1 = load_word(p + 4), w2 = load_word(p + 8);
    uint4 px;
    widen_rgb24(w0, w1, w2, px.x, px.y, px.z, px.w);   // block_codec.cuh
    return px;
}

template <int MODE, int FLAVOUR>
__global__ void __launch_bounds__(256, rgb24_ctas_per_sm(MODE)) encode_rgb24_kernel(const EncodeParams P)
{
    __shared__ uint32_t lut[MODE == kDxt1 ? 1 : 256];
    pdl_launch_dependents();
    if (MODE != kDxt1) stage_control_lut<FLAVOUR != 0, 256>(lut, threadIdx.y * blockDim.x + threadIdx.x);
    const uint32_t bx = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t by = blockIdx.y * blockDim.y + threadIdx.y;
    const uint8_t* src = P.src + (uint64_t)blockIdx.z * P.srcPitch + (uint64_t)bx * 12u;
    uint8_t* dst = P.dst + (uint64_t)blockIdx.z * P.dstPitch + (uint64_t)bx * 8u;
    uint8_t* dst2 = MODE == kDual ? P.dst2 + (uint64_t)blockIdx.z * P.dstPitch + (uint64_t)bx * 8u : nullptr;
    pdl_wait();
    if (MODE != kDxt1) __syncthreads();
    if (bx >= P.bw || by >= P.bh) return;
    const uint32_t rowStep = gridDim.y * blockDim.y;
#pragma unroll 1
    for (; by < P.bh; by += rowStep) {
        const uint8_t* s = src + (uint64_t)by * 4u * P.stride;
        const uint4 r0 = load_row_rgb24(s);
        const uint4 r1 = load_row_rgb24(s + P.stride);
        const uint4 r2 = load_row_rgb24(s + 2ull * P.stride);
        const uint4 r3 = load_row_rgb24(s + 3ull * P.stride);
        const uint64_t o = (uint64_t)by * P.bw * 8u;
        encode_and_store_flavour<MODE, FLAVOUR>(r0, r1, r2, r3, lut, dst + o, MODE == kDual ? dst2 + o : nullptr);
    }
}

// Relaxed shapes: any width / height >= 1 and any 4-byte-aligned stride.  Blocks that
// hang over the right or bottom edge replicate the last column / row (clamp-to-edge), so the output has
// ceil(w/4) x ceil(h/4) blocks.  Blocks that lie wholly inside the image take the four 128-bit row loads of the
// strict kernels when the rows are 16-byte aligned (a uniform condition); edge blocks and unaligned images take
// sixteen clamped 32-bit loads.  The DXT1 flavours are launched as one-shot CTAs (grid.y = block rows), the ETC1s
// flavours with CTAs that walk four block rows each.  FLAVOUR 0 = SSE2-exact arithmetic, 1 = float-reference arithmetic; on images the
// strict entry points accept, the bytes are the same.
template <int CODEC, int FLAVOUR>
__global__ void __launch_bounds__(256, CODEC == kDxt1 ? 6 : GB_RELAXED_ETC1_CTAS) encode_relaxed_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                                  uint32_t width, uint32_t height, uint32_t stride)
{
    constexpr bool ROWS = CODEC != kDxt1;   // DXT1 flavours: one-shot CTAs (gridDim.y = block rows)
    __shared__ uint32_t lut[CODEC == kDxt1 ? 1 : 256];
    pdl_launch_dependents();
    if (CODEC != kDxt1) stage_control_lut<FLAVOUR != 0, 256>(lut, threadIdx.x);
    pdl_wait();
    if (CODEC != kDxt1) __syncthreads();
    const uint32_t bw = (width + 3u) / 4u, bh = (height + 3u) / 4u;
    const uint32_t bx = blockIdx.x * 256u + threadIdx.x;
    if (bx >= bw) return;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | stride) & 15u) == 0u;
    const bool insideX = bx * 4u + 3u < width;
    // CTAs walk down the image (gridDim.y apart): the table staging and the column tests are paid once per few blocks
    uint32_t by = blockIdx.y;
    if (by >= bh) return;
#pragma unroll 1
    do {
        uint32_t p[16];
        if (aligned && insideX && by * 4u + 3u < height) {
            const uint8_t* s = src + (uint64_t)by * 4u * stride + (uint64_t)bx * 16u;
            const uint4 r0 = load_row(s), r1 = load_row(s + stride), r2 = load_row(s + 2ull * stride), r3 = load_row(s + 3ull * stride);
            p[0] = r0.x; p[1] = r0.y; p[2] = r0.z; p[3] = r0.w;
            p[4] = r1.x; p[5] = r1.y; p[6] = r1.z; p[7] = r1.w;
            p[8] = r2.x; p[9] = r2.y; p[10] = r2.z; p[11] = r2.w;
            p[12] = r3.x; p[13] = r3.y; p[14] = r3.z; p[15] = r3.w;
        } else {
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const uint32_t row = min(by * 4u + (uint32_t)y, height - 1u);
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const uint32_t col = min(bx * 4u + (uint32_t)x, width - 1u);
                    p[4 * y + x] = __ldg(reinterpret_cast<const uint32_t*>(src + (uint64_t)row * stride + (uint64_t)col * 4u));
                }
            }
        }
        uint32_t w0, w1;
        if (FLAVOUR == 0) {
            const BlockFront f = analyse(p);
            if (CODEC == kDxt1) encode_dxt1<GB_SEL_DXT1>(p, f, w0, w1);
            else encode_etc1(p, f, lut, w0, w1);
        } else {
            const RefFront f = analyse_ref(p, CODEC == kDxt1 ? 32u : 64u);
            if (CODEC == kDxt1) encode_dxt1_ref(p, f, w0, w1);
            else encode_etc1_ref(p, f, lut, w0, w1);
        }
        store_block(dst + ((uint64_t)by * bw + bx) * 8u, w0, w1);
    } while (ROWS && (by += gridDim.y) < bh);
}

// ------------------------------------------------------------------ ragged batches
// Images of different shapes in one launch (mip chains, atlases).  The host sorts nothing: it provides the
// descriptors plus an exclusive prefix sum of per-image CTA counts; each CTA finds its image by binary search
// and its tile by one division.  A tile is 64 x 16 blocks, walked in four passes of 64 x 4 threads, so the
// search and (ETC1s) the control-table staging are paid once per 1024 blocks.
// Small batches (a mip chain is 10-14 images) travel as kernel parameters: the search then runs against the
// constant bank instead of a chain of dependent global loads before the first pixel load can be issued.
struct BatchImage {
    const uint8_t* src;
    uint8_t* dst;
    uint8_t* dst2;    // dual-output batches: the ETC1s blocks
    uint32_t bw, bh;
    uint32_t stride;
    uint32_t tilesX;  // CTAs per block row
};

constexpr int kBatchTileX = 64;    // blocks per CTA in x
constexpr int kBatchTileY = 4;     // block rows per pass (threads in y)
constexpr int kBatchPasses = 4;    // passes per CTA
constexpr int kBatchInline = 40;   // largest batch passed as kernel parameters (40 x 44 bytes: well inside the 4 KiB limit)

struct BatchTableGlobal {
    const BatchImage* images;
    const uint32_t* ctaStart;
    __device__ __forceinline__ uint32_t start(uint32_t i) const { return ctaStart[i]; }
    __device__ __forceinline__ BatchImage image(uint32_t i) const { return images[i]; }
};
struct BatchTableInline {
    BatchImage images[kBatchInline];
    uint32_t ctaStart[kBatchInline];
    __device__ __forceinline__ uint32_t start(uint32_t i) const { return ctaStart[i]; }
    __device__ __forceinline__ BatchImage image(uint32_t i) const { return images[i]; }
};

template <int MODE, int FLAVOUR, typename TABLE>
__global__ void __launch_bounds__(kBatchTileX* kBatchTileY, ctas_per_sm(MODE))
    encode_batch_kernel(const __grid_constant__ TABLE table, uint32_t nImages)
{
    __shared__ uint32_t lut[MODE == kDxt1 ? 1 : 256];
    pdl_launch_dependents();
    if (MODE != kDxt1) stage_control_lut<FLAVOUR != 0, kBatchTileX * kBatchTileY>(lut, threadIdx.y * kBatchTileX + threadIdx.x);
    pdl_wait();   // (the descriptor table of a large batch is global memory written by a copy earlier in the stream)
    if (MODE != kDxt1) __syncthreads();
    // largest i with start(i) <= blockIdx.x
    uint32_t lo = 0, hi = nImages;
    while (hi - lo > 1u) {
        const uint32_t m = (lo + hi) >> 1;
        if (table.start(m) <= blockIdx.x) lo = m; else hi = m;
    }
    const BatchImage im = table.image(lo);
    const uint32_t local = blockIdx.x - table.start(lo);
    const uint32_t ty = local / im.tilesX, tx = local - ty * im.tilesX;
    const uint32_t bx = tx * kBatchTileX + threadIdx.x;
    if (bx >= im.bw) return;
#pragma unroll 1
    for (uint32_t pass = 0; pass < (uint32_t)kBatchPasses; ++pass) {
        const uint32_t by = (ty * kBatchPasses + pass) * kBatchTileY + threadIdx.y;
        if (by >= im.bh) return;
        const uint8_t* s = im.src + (uint64_t)by * 4u * im.stride + (uint64_t)bx * 16u;
        const uint4 r0 = load_row(s);
        const uint4 r1 = load_row(s + im.stride);
        const uint4 r2 = load_row(s + 2ull * im.stride);
        const uint4 r3 = load_row(s + 3ull * im.stride);
        const uint64_t o = ((uint64_t)by * im.bw + bx) * 8u;
        encode_and_store_flavour<MODE, FLAVOUR>(r0, r1, r2, r3, lut, im.dst + o, MODE == kDual ? im.dst2 + o : nullptr);
    }
}

}  // namespace gb
