// tma_kernels.cuh -- second image load/tiling layer: TMA 2D tiles staged in shared memory.
//
// Same per-block arithmetic as encode_direct_kernel (block_codec.cuh), different way of getting the 16 pixels
// into registers.  A CTA is one PRODUCER warp and 2*RB consumer warps around a ring of shared-memory stages:
//   producer   one lane walks the CTA's tiles and asks the TMA unit (cp.async.bulk.tensor, SASS UTMALDG) for one
//              box per tile -- 256 pixels x 4*RB pixel rows = 64 x RB blocks, 4*RB KiB -- as soon as the stage's
//              "empty" mbarrier says all consumer warps have read the previous tile out of it;
//   consumers  one thread per block of the tile: wait on the stage's "full" mbarrier, four conflict-free LDS.128,
//              one arrive per WARP on the "empty" mbarrier (after a __syncwarp, not a CTA barrier), encode, store.
// Nothing in the loop synchronises the CTA: consumer warps drift up to a ring apart, and the bytes in flight per
// SM (CTAs x (stages - 1) x box) do not depend on which phase the warps are in.
// Box size is what the TMA unit is sensitive to (measured on B200, profiles/r02_shape_ab.md): warp-private rings
// of 2 KiB boxes reach 2.6-4.2 TB/s, 4 KiB boxes 6.0-6.2 TB/s (about 190 cycles per box per SM whatever its size),
// so tiles are 8-16 KiB.  (Round-1 design, in the git history: 4 KiB boxes, a CTA-wide __syncthreads per tile.)
//
// The tensor map carries the row stride and the image pitch, so padded strides and uniform batches cost no address
// arithmetic in the kernel, and out-of-range columns / rows of edge tiles are zero-filled by the hardware.
//
// Replaces the same reference lines as encode_kernels.cuh (GoofyTC/goofy_tc.h:1077-1099 tile fetch, :1514-1524
// loops); used for strided / batched inputs when the launcher's policy picks it (host_launch.cuh: choose_tma).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "encode_kernels.cuh"

namespace gb {

constexpr int kTmaMaxStages = 8;
constexpr int kTmaBoxPixels = 256;   // TMA box limit: 256 elements per dimension (u32 element = 1 pixel) = 64 blocks
constexpr int kTmaTileBlocksX = kTmaBoxPixels / 4;
constexpr int kTmaRowBytes = kTmaBoxPixels * 4;       // 1 KiB per pixel row of a box
constexpr int kTmaBlockRowBytes = 4 * kTmaRowBytes;   // 4 KiB: the four pixel rows of 64 blocks
constexpr int tma_threads(int rb) { return (2 * rb + 1) * 32; }   // 2*RB consumer warps + the producer warp
// resident CTAs per SM the kernel is compiled for: 32 consumer warps per SM whatever the tile height
#ifndef GB_TMA_CONSUMER_WARPS_PER_SM
#define GB_TMA_CONSUMER_WARPS_PER_SM 32
#endif
constexpr int tma_min_ctas(int rb) { return GB_TMA_CONSUMER_WARPS_PER_SM / (2 * rb); }

// q = n / d by multiplication: exact for n < 2^24 and d <= 2^16 (m = ceil(2^40 / d)).
struct FastDiv {
    uint64_t m;
    uint32_t d;
    __host__ __device__ uint32_t div(uint32_t n) const { return (uint32_t)(((uint64_t)n * m) >> 40); }
};

struct TmaParams {
    uint8_t* dst;
    uint8_t* dst2;
    uint64_t dstPitch;
    uint32_t bw, bh;
    uint32_t nTiles;      // tilesX x rowGroups x images
    uint32_t nStages;     // depth of the CTA's ring (dynamic smem = nStages x RB x 4 KiB)
    uint32_t evictFirst;  // 1: loads carry an L2 evict-first policy (streaming data)
    FastDiv tilesX;       // tiles per row group (ceil(bw / 64))
    FastDiv groups;       // row groups per image (ceil(bh / RB))
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// box (c0 pixels, c1 pixel rows, c2 images) -> shared memory, optionally with an L2 evict-first policy
__device__ __forceinline__ void tma_load_box(uint32_t dstSmem, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t c2,
                                             uint32_t bar, uint64_t policy, bool hinted)
{
    if (hinted)
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
            " [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(dstSmem),
            "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy)
            : "memory");
    else
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dstSmem),
            "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
            : "memory");
}
// the same box into L2 only (no shared memory, no barrier): harmless before griddepcontrol.wait
__device__ __forceinline__ void tma_prefetch_box(const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// RB = block rows per tile (box = 256 pixels x 4*RB rows).  blockDim.x = tma_threads(RB).
template <int MODE, int RB>
__global__ void __launch_bounds__(tma_threads(RB), tma_min_ctas(RB)) encode_tma_kernel(const __grid_constant__ CUtensorMap map, const TmaParams P)
{
    extern __shared__ __align__(1024) uint8_t ringMem[];   // [stage][4*RB pixel rows][1 KiB]
    __shared__ __align__(8) uint64_t fullBar[kTmaMaxStages];
    __shared__ __align__(8) uint64_t emptyBar[kTmaMaxStages];
    __shared__ uint32_t lut[MODE == kDxt1 ? 1 : 256];
    constexpr uint32_t kStageBytes = (uint32_t)RB * kTmaBlockRowBytes;
    constexpr uint32_t kConsumers = 2u * RB * 32u;

    pdl_launch_dependents();
    const uint32_t tid = threadIdx.x;
    if (MODE != kDxt1 && tid < kConsumers) stage_control_lut<false, (int)kConsumers>(lut, tid);
    if (tid == kConsumers) {
        for (uint32_t s = 0; s < P.nStages; ++s) {
            mbar_init(smem_addr(&fullBar[s]), 1);
            mbar_init(smem_addr(&emptyBar[s]), 2u * RB);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();   // barriers and control table are in place; the only CTA-wide barrier of the kernel

    // tiles are dealt round-robin over the CTAs of the grid, x fastest: at any moment the whole chip works on one
    // band of pixel rows, and the per-CTA tile counts differ by at most one
    const uint32_t G = gridDim.x;
    const uint32_t nMine = blockIdx.x < P.nTiles ? (P.nTiles - 1u - blockIdx.x) / G + 1u : 0u;
    auto coords = [&](uint32_t tile, uint32_t& tx, uint32_t& rg, uint32_t& img) {
        const uint32_t r = P.tilesX.div(tile);
        tx = tile - r * P.tilesX.d;
        img = P.groups.div(r);
        rg = r - img * P.groups.d;
    };

    if (tid >= kConsumers) {
        // ------------------------------------------------------------ producer warp (one lane works)
        if (tid != kConsumers) return;
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
        uint64_t policy = 0;
        if (P.evictFirst) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        // before the previous kernel of the stream has finished, the first tiles may only be pulled into L2
#pragma unroll 1
        for (uint32_t s = 0; s < P.nStages && s < nMine; ++s) {
            uint32_t tx, rg, img;
            coords(blockIdx.x + s * G, tx, rg, img);
            tma_prefetch_box(&map, tx * (uint32_t)kTmaBoxPixels, rg * (uint32_t)(4 * RB), img);
        }
        pdl_wait();
        const uint32_t ringBase = smem_addr(ringMem);
        uint32_t stage = 0, parity = 1;   // a fresh mbarrier passes a wait on parity 1: the ring starts out empty
#pragma unroll 1
        for (uint32_t it = 0; it < nMine; ++it) {
            uint32_t tx, rg, img;
            coords(blockIdx.x + it * G, tx, rg, img);
            while (!mbar_try_wait(smem_addr(&emptyBar[stage]), parity)) {}
            const uint32_t bar = smem_addr(&fullBar[stage]);
            mbar_arrive_expect_tx(bar, kStageBytes);   // a box always delivers its full size (out-of-range parts arrive as zeros)
            tma_load_box(ringBase + stage * kStageBytes, &map, tx * (uint32_t)kTmaBoxPixels, rg * (uint32_t)(4 * RB), img, bar, policy,
                         P.evictFirst != 0u);
            if (++stage == P.nStages) { stage = 0; parity ^= 1u; }
        }
        return;
    }

    // ---------------------------------------------------------------- consumer warps: one thread per block of the tile
    const uint32_t col = tid & 63u, row = tid >> 6;
    const uint8_t* mine = ringMem + row * kTmaBlockRowBytes + col * 16u;
    uint32_t stage = 0, parity = 0;
#pragma unroll 1
    for (uint32_t it = 0; it < nMine; ++it) {
        uint32_t tx, rg, img;
        coords(blockIdx.x + it * G, tx, rg, img);
        const uint32_t bx = tx * (uint32_t)kTmaTileBlocksX + col, by = rg * (uint32_t)RB + row;
        while (!mbar_try_wait(smem_addr(&fullBar[stage]), parity)) {}
        const uint8_t* sm = mine + stage * kStageBytes;
        const uint4 r0 = *reinterpret_cast<const uint4*>(sm);
        const uint4 r1 = *reinterpret_cast<const uint4*>(sm + kTmaRowBytes);
        const uint4 r2 = *reinterpret_cast<const uint4*>(sm + 2 * kTmaRowBytes);
        const uint4 r3 = *reinterpret_cast<const uint4*>(sm + 3 * kTmaRowBytes);
        // every lane of this warp holds its block in registers: the warp is done with the stage
        __syncwarp();
        if ((tid & 31u) == 0u) mbar_arrive(smem_addr(&emptyBar[stage]));
        if (++stage == P.nStages) { stage = 0; parity ^= 1u; }
        if (bx < P.bw && by < P.bh) {
            const uint64_t o = (uint64_t)img * P.dstPitch + ((uint64_t)by * P.bw + bx) * 8u;
            encode_and_store<MODE>(r0, r1, r2, r3, lut, P.dst + o, MODE == kDual ? P.dst2 + o : nullptr);
        }
    }
}

}  // namespace gb
