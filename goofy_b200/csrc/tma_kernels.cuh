// tma_kernels.cuh -- second image load/tiling layer: TMA 2D tiles staged in shared memory.
//
// Same per-block arithmetic as encode_direct_kernel (block_codec.cuh), different way of getting
// the 16 pixels into registers: persistent CTAs walk (image, block row, x tile) tiles; one
// elected thread asks the TMA unit for the tile's pixel rows (cp.async.bulk.tensor, SASS
// UTMALDG) into a 4-stage shared-memory ring guarded by mbarriers; every thread then reads its
// block's four rows with conflict-free LDS.128.  The tensor map carries the row stride and the
// image pitch, so padded strides and uniform batches cost no address arithmetic in the kernel,
// and out-of-range columns of the last tile are zero-filled by the hardware.
//
// Replaces the same reference lines as encode_kernels.cuh (GoofyTC/goofy_tc.h:1077-1099 tile
// fetch, :1514-1524 loops); used for strided / batched inputs when the launcher's policy picks
// it (capi.cu: choose_load_path).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "encode_kernels.cuh"

namespace gb {

constexpr int kTmaMaxStages = 8;
constexpr int kTmaBoxPixels = 256;   // TMA box limit: 256 elements per dimension (u32 element = 1 pixel)
#ifndef GB_TMA_BOXES
#define GB_TMA_BOXES 2
#endif
constexpr int kTmaBoxesPerTile = GB_TMA_BOXES;  // tile = 512 pixels = 128 blocks wide (128-thread CTAs), one block row high
constexpr int kTmaThreads = kTmaBoxesPerTile * kTmaBoxPixels / 4;             // 128: one thread per block
constexpr int kTmaBoxBytes = kTmaBoxPixels * 4 * 4;                           // 4 pixel rows x 1 KiB
constexpr int kTmaStageBytes = kTmaBoxesPerTile * kTmaBoxBytes;               // 8 KiB

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
    uint32_t nTiles;
    uint32_t nStages;  // depth of the shared-memory tile ring (dynamic smem = nStages * 16 KiB)
    uint32_t evictFirst;  // 1: loads carry an L2 evict-first policy (streaming data)
    FastDiv tilesX;  // tiles per block row
    FastDiv rows;    // block rows per image
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

template <int MODE>
__global__ void __launch_bounds__(kTmaThreads) encode_tma_kernel(const __grid_constant__ CUtensorMap map, const TmaParams P)
{
    extern __shared__ __align__(1024) uint8_t tileMem[];
    __shared__ __align__(8) uint64_t fullBar[kTmaMaxStages];
    __shared__ uint32_t lut[MODE == kDxt1 ? 1 : 256];

    const uint32_t tid = threadIdx.x;
    if (MODE != kDxt1) {
#pragma unroll
        for (uint32_t i = tid; i < 256u; i += kTmaThreads) lut[i] = g_etc1ControlLut[i];
    }
    if (tid == 0) {
        for (uint32_t s = 0; s < P.nStages; ++s) mbar_init(smem_addr(&fullBar[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t G = gridDim.x;
    const uint32_t nMine = blockIdx.x < P.nTiles ? (P.nTiles - 1u - blockIdx.x) / G + 1u : 0u;
    const uint32_t tileBase = smem_addr(tileMem);

    uint64_t policy = 0;
    auto issue = [&](uint32_t tile, uint32_t stage) {
        const uint32_t r = P.tilesX.div(tile), tx = tile - r * P.tilesX.d;
        const uint32_t img = P.rows.div(r), by = r - img * P.rows.d;
        const uint32_t blocksLeft = P.bw - tx * (kTmaThreads);
        const uint32_t nBoxes = blocksLeft >= (uint32_t)kTmaThreads ? (uint32_t)kTmaBoxesPerTile : (blocksLeft + 63u) / 64u;
        const uint32_t bar = smem_addr(&fullBar[stage]);
        mbar_arrive_expect_tx(bar, nBoxes * kTmaBoxBytes);
#pragma unroll 1
        for (uint32_t j = 0; j < nBoxes; ++j)
            tma_load_box(tileBase + stage * kTmaStageBytes + j * kTmaBoxBytes, &map,
                         (tx * kTmaBoxesPerTile + j) * kTmaBoxPixels, by * 4u, img, bar, policy, P.evictFirst != 0u);
    };
    if (tid == 0) {
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
#pragma unroll 1
        for (uint32_t s = 0; s < P.nStages && s < nMine; ++s) issue(blockIdx.x + s * G, s);
    }

    uint32_t stage = 0, parity = 0;
#pragma unroll 1
    for (uint32_t it = 0; it < nMine; ++it) {
        const uint32_t tile = blockIdx.x + it * G;
        const uint32_t r = P.tilesX.div(tile), tx = tile - r * P.tilesX.d;
        const uint32_t img = P.rows.div(r), by = r - img * P.rows.d;

        const uint32_t bar = smem_addr(&fullBar[stage]);
        while (!mbar_try_wait(bar, parity)) {}
        const uint8_t* sm = tileMem + stage * kTmaStageBytes + (tid >> 6) * kTmaBoxBytes + (tid & 63u) * 16u;
        const uint4 r0 = *reinterpret_cast<const uint4*>(sm);
        const uint4 r1 = *reinterpret_cast<const uint4*>(sm + kTmaBoxPixels * 4);
        const uint4 r2 = *reinterpret_cast<const uint4*>(sm + 2 * kTmaBoxPixels * 4);
        const uint4 r3 = *reinterpret_cast<const uint4*>(sm + 3 * kTmaBoxPixels * 4);
        __syncthreads();  // every thread holds its block in registers: the stage can be refilled
        if (tid == 0 && it + P.nStages < nMine) issue(tile + P.nStages * G, stage);
        const uint32_t stageDone = stage;
        if (++stage == P.nStages) { stage = 0; parity ^= 1u; }
        (void)stageDone;

        const uint32_t bx = tx * kTmaThreads + tid;
        if (bx < P.bw) {
            const uint64_t o = (uint64_t)img * P.dstPitch + ((uint64_t)by * P.bw + bx) * 8u;
            encode_and_store<MODE>(r0, r1, r2, r3, lut, P.dst + o, MODE == kDual ? P.dst2 + o : nullptr);
        }
    }
}

}  // namespace gb
