// host_launch.cuh -- launch configuration of the device-resident entry points: load-layer policy (one-shot,
// row-walking, cp.async ring, TMA tiles), programmatic dependent launch, the float-reference flavour.
#pragma once
#include "host_common.cuh"

namespace {

template <typename Kernel>
int launch_encode(Kernel kernel, dim3 grid, dim3 block, cudaStream_t stream, const gb::EncodeParams& P, const char* name)
{
    t_lastKernel = name;
    static const bool pdl = []() { const char* e = getenv("GOOFY_B200_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, P);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(e);
}

// Any kernel of this library with programmatic stream serialisation: every one of them starts with
// griddepcontrol.launch_dependents and waits (griddepcontrol.wait) before it touches global memory.
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args&&... args)
{
    static const bool pdl = []() { const char* e = getenv("GOOFY_B200_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(e);
}

template <int MODE>
constexpr const char* kernel_name(const char* dxt1, const char* etc1, const char* dual) { return MODE == gb::kDxt1 ? dxt1 : MODE == gb::kEtc1 ? etc1 : dual; }
#define GB_KNAME(base) kernel_name<MODE>(base "<dxt1>", base "<etc1s>", base "<dxt1+etc1s>")

template <int MODE, bool PITCHED>
int launch_direct_grid(const gb::EncodeParams& Q, dim3 grid, dim3 block, cudaStream_t stream)
{
    // 32-bit in-image offsets unless the image spans 4 GiB or more
    const bool narrow = (uint64_t)Q.bh * 4u * Q.stride + (uint64_t)Q.bw * 16u < 0xFFFFFFFFull;
    if (Q.firstWave != 0u)   // short launch: the instantiation with the L2 warm-up
        return narrow ? launch_encode(gb::encode_direct_kernel<MODE, false, PITCHED, true>, grid, block, stream, Q, GB_KNAME("encode_direct_kernel[short launch]"))
                      : launch_encode(gb::encode_direct_kernel<MODE, true, PITCHED, true>, grid, block, stream, Q, GB_KNAME("encode_direct_kernel[short launch]"));
    return narrow ? launch_encode(gb::encode_direct_kernel<MODE, false, PITCHED, false>, grid, block, stream, Q, GB_KNAME("encode_direct_kernel"))
                  : launch_encode(gb::encode_direct_kernel<MODE, true, PITCHED, false>, grid, block, stream, Q, GB_KNAME("encode_direct_kernel"));
}

template <int MODE, bool PITCHED>
int launch_rows_grid(const gb::EncodeParams& Q, dim3 grid, dim3 block, cudaStream_t stream, bool narrow)
{
    if (Q.firstWave != 0u || Q.prefetchNext != 0u)
        return narrow ? launch_encode(gb::encode_rows_kernel<MODE, false, PITCHED, true>, grid, block, stream, Q, GB_KNAME("encode_rows_kernel[short launch]"))
                      : launch_encode(gb::encode_rows_kernel<MODE, true, PITCHED, true>, grid, block, stream, Q, GB_KNAME("encode_rows_kernel[short launch]"));
    return narrow ? launch_encode(gb::encode_rows_kernel<MODE, false, PITCHED, false>, grid, block, stream, Q, GB_KNAME("encode_rows_kernel"))
                  : launch_encode(gb::encode_rows_kernel<MODE, true, PITCHED, false>, grid, block, stream, Q, GB_KNAME("encode_rows_kernel"));
}

int sm_count(int dev);
int env_int(const char* name, int lo, int hi, int fallback);

// CTAs of a launch that warm L2 with their first tile before griddepcontrol.wait (encode_kernels.cuh: prefetch_l2):
// the ones resident at once.  GOOFY_B200_L2PF=0 turns it off.
uint32_t first_wave(uint32_t resident)
{
    static const bool on = env_int("GOOFY_B200_L2PF", 0, 1, 1) != 0;
    return on ? resident : 0u;
}

// The L2 warm-up before griddepcontrol.wait pays on short launches (a 16384 x 2048 strip: +4-5 %, profiles/r02_shape_ab.md),
// where the hand-over between two launches is a visible share of the time; long launches skip it.
constexpr uint32_t kL2PrefetchMaxWaves = 16;

// Row-walking launch: one (possibly very tall) image, or a batch at fixed pitches (grid.z = image).
template <int MODE>
int launch_rows(const gb::EncodeParams& P, uint32_t nImages, cudaStream_t stream, int dev)
{
    uint32_t tx = 32u;
    while (tx < (uint32_t)GB_TPB && tx < P.bw) tx <<= 1;
    const uint32_t ty = (uint32_t)GB_TPB / tx;
    const uint32_t gx = (P.bw + tx - 1u) / tx;
    const uint32_t rowGroups = (P.bh + ty - 1u) / ty;
    const int sms = sm_count(dev);
    if (sms <= 0) return GOOFY_B200_E_DEVICE;
    // AUTO: the dual-output kernel walks its rows through the cp.async ring (next block in flight while the current one
    // is encoded): +2-3 % over plain loads in every A/B session since it stopped being ALU-bound (r01f sessions 3, 8, 9);
    // the single-codec kernels are faster with plain loads (ETC1s -5 % through the ring)
    const int loadPath = g_loadPath.load(std::memory_order_relaxed);
    // ... on long launches: a short one (a 16384 x 2048 strip) is 9 % faster with plain loads (5576 vs 6089 GB/s, session D)
    const uint32_t rowsPerCtaNominal = MODE == gb::kDual ? 8u : 4u;
    const bool longLaunch = (uint64_t)gx * ((rowGroups + rowsPerCtaNominal - 1u) / rowsPerCtaNominal) >=
                            (uint64_t)kL2PrefetchMaxWaves * (uint32_t)sms * (uint32_t)gb::ctas_per_sm(MODE);
    const bool async = nImages == 1u && (loadPath == GOOFY_B200_LOAD_ASYNC || (loadPath == GOOFY_B200_LOAD_AUTO && MODE == gb::kDual && longLaunch));
    const uint32_t resident = async ? (uint32_t)sms * (uint32_t)GB_ASYNC_CTAS(MODE)
                                    : (uint32_t)sms * (uint32_t)gb::ctas_per_sm(MODE) * (256u / (uint32_t)GB_TPB);
    // Each CTA walks a few block rows: enough to amortise the per-thread set-up, few enough that CTAs keep
    // retiring and restarting at staggered times (fully persistent CTAs run in lock-step and are 10 % slower;
    // profiles/r01_rows_grid_sweep.txt).  Never fewer CTAs than one resident wave.
    // Measured on the batched 4 x 8192^2 launch: ETC1s 6915 / 7011 / 7010 / 6966 / 6892 GB/s at 2 / 3 / 4 / 6 / 8 rows
    // per CTA; dual-output 5835 / 5932 / 6055 / 6193 / 6293.
    static const uint32_t rowsEnv = (uint32_t)env_int("GOOFY_B200_ROWS_PER_CTA", 1, 1 << 20, 0);
    const uint32_t rowsPerCta = rowsEnv ? rowsEnv : (MODE == gb::kDual ? 8u : 4u);
    uint32_t gy = (rowGroups + rowsPerCta - 1u) / rowsPerCta;
    // Short launches: a grid of, say, 1.4 resident waves leaves the chip half empty for the second one.  Size the grid to
    // a whole number of waves instead (the nearest, at least one) and let the rows per CTA come out as they may.
    static const bool quantise = env_int("GOOFY_B200_WAVE_QUANT", 0, 1, 1) != 0;
    const uint64_t ctas0 = (uint64_t)gx * gy * nImages;
    if (quantise && ctas0 < (uint64_t)kL2PrefetchMaxWaves * resident) {
        uint64_t waves = (ctas0 + resident / 2u) / resident;
        if (waves == 0u) waves = 1u;
        gy = (uint32_t)(waves * resident / ((uint64_t)gx * nImages));
    } else {
        const uint32_t perImage = (resident + nImages - 1u) / nImages;   // this image's share of one resident wave
        if (gy < perImage / gx) gy = perImage / gx;
    }
    if (gy == 0u) gy = 1u;
    if (gy > rowGroups) gy = rowGroups;
    if (gy > 65535u) gy = 65535u;
    const dim3 block(tx, ty, 1);
    const bool narrow = (uint64_t)P.bh * 4u * P.stride + (uint64_t)P.bw * 16u < 0xFFFFFFFFull;
    for (uint32_t img0 = 0; img0 < nImages; img0 += 65535u) {
        const uint32_t nz = nImages - img0 < 65535u ? nImages - img0 : 65535u;
        gb::EncodeParams Q = P;
        Q.src += (uint64_t)img0 * P.srcPitch;
        Q.dst += (uint64_t)img0 * P.dstPitch;
        if (Q.dst2) Q.dst2 += (uint64_t)img0 * P.dstPitch;
        const dim3 grid(gx, gy, nz);
        Q.firstWave = (!async && (uint64_t)gx * gy * nz <= (uint64_t)kL2PrefetchMaxWaves * resident) ? first_wave(resident) : 0u;
        static const bool pfNext = env_int("GOOFY_B200_PF_NEXT", 0, 1, 1) != 0;   // +2-3 % on short ETC1s / dual launches (session E)
        Q.prefetchNext = (pfNext && Q.firstWave != 0u) ? 1u : 0u;
        int rc;
        if (async)
            rc = narrow ? launch_encode(gb::encode_rows_async_kernel<MODE, false>, grid, block, stream, Q, GB_KNAME("encode_rows_async_kernel"))
                        : launch_encode(gb::encode_rows_async_kernel<MODE, true>, grid, block, stream, Q, GB_KNAME("encode_rows_async_kernel"));
        else if (nImages > 1u)
            rc = launch_rows_grid<MODE, true>(Q, grid, block, stream, narrow);
        else
            rc = launch_rows_grid<MODE, false>(Q, grid, block, stream, narrow);
        if (rc != GOOFY_B200_OK) return rc;
    }
    return GOOFY_B200_OK;
}

template <int MODE>
int launch_direct(gb::EncodeParams P, uint32_t nImages, cudaStream_t stream, int dev)
{
    // A batch whose images lie back to back (pitch == image size) is one tall image.
    const uint64_t imageBytes = (uint64_t)P.bh * 4u * P.stride, outBytes = (uint64_t)P.bh * P.bw * 8u;
    if (nImages > 1u && P.srcPitch == imageBytes && P.dstPitch == outBytes && (uint64_t)P.bh * nImages <= 0xFFFFFFFFull) {
        P.bh *= nImages;
        nImages = 1u;
    }
    // Load-path policy for AUTO: the DXT1 kernel is HBM-bound either way and
    // one-shot CTAs are marginally faster (6617 vs 6598 GB/s); the ETC1s and dual-output kernels
    // gain 4-13 % from row-walking CTAs (dual-output on a single image: through the cp.async ring, see launch_rows).
    const int path = g_loadPath.load(std::memory_order_relaxed);
    const bool rows = path == GOOFY_B200_LOAD_DIRECT || path == GOOFY_B200_LOAD_ASYNC ||
                      (path == GOOFY_B200_LOAD_AUTO && MODE != gb::kDxt1);
    if (rows) return launch_rows<MODE>(P, nImages, stream, dev);
    // one-shot CTAs: threads: x walks blocks along a row (coalescing), y stacks block rows for narrow images
    // (x is a power of two and x*y == GB_TPB: the kernels rely on exactly GB_TPB threads)
    uint32_t tx = 32u;
    while (tx < (uint32_t)GB_TPB && tx < P.bw) tx <<= 1;
    const uint32_t ty = (uint32_t)GB_TPB / tx;
    const dim3 block(tx, ty, 1);
    const uint32_t gx = (P.bw + tx - 1u) / tx;
    const uint32_t rowsPerLaunch = 65535u * ty;
    const int sms = sm_count(dev);
    const uint32_t resident = (uint32_t)(sms > 0 ? sms : 148) * (uint32_t)gb::ctas_per_sm(MODE) * (256u / (uint32_t)GB_TPB);
    const uint64_t totalCtas = (uint64_t)gx * ((P.bh + ty - 1u) / ty) * nImages;
    P.firstWave = totalCtas <= (uint64_t)kL2PrefetchMaxWaves * resident ? first_wave(resident) : 0u;
    for (uint32_t img0 = 0; img0 < nImages; img0 += 65535u) {
        const uint32_t nz = nImages - img0 < 65535u ? nImages - img0 : 65535u;
        for (uint32_t by0 = 0; by0 < P.bh; by0 += rowsPerLaunch) {
            const uint32_t rows = P.bh - by0 < rowsPerLaunch ? P.bh - by0 : rowsPerLaunch;
            gb::EncodeParams Q = P;
            Q.by0 = by0;
            Q.src += (uint64_t)img0 * P.srcPitch;
            Q.dst += (uint64_t)img0 * P.dstPitch;
            if (Q.dst2) Q.dst2 += (uint64_t)img0 * P.dstPitch;
            const dim3 grid(gx, (rows + ty - 1u) / ty, nz);
            const int rc = nImages > 1u ? launch_direct_grid<MODE, true>(Q, grid, block, stream)
                                        : launch_direct_grid<MODE, false>(Q, grid, block, stream);
            if (rc != GOOFY_B200_OK) return rc;
        }
    }
    return GOOFY_B200_OK;
}

// ---------------------------------------------------------------- TMA tile path

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

struct DeviceInfo {
    int smCount = 0;
    int tmaCtasPerSm[3][3] = {};  // [mode][log2 RB]; 0 = not yet configured
    int tmaSmemBytes[3][3] = {};
};
DeviceInfo g_devInfo[kMaxDevices];
std::mutex g_devInfoMutex;

int sm_count(int dev)
{
    std::lock_guard<std::mutex> g(g_devInfoMutex);
    DeviceInfo& di = g_devInfo[dev];
    if (di.smCount == 0 && cudaDeviceGetAttribute(&di.smCount, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        cudaGetLastError();
        di.smCount = 0;
    }
    return di.smCount;
}

int env_int(const char* name, int lo, int hi, int fallback)
{
    const char* e = getenv(name);
    if (!e || !*e) return fallback;
    const int v = atoi(e);
    return (v >= lo && v <= hi) ? v : fallback;
}

// Depth of the tile ring and block rows per tile.  GOOFY_B200_TMA_STAGES (2..8) and GOOFY_B200_TMA_ROWS (1, 2, 4)
// override them for experiments.
uint32_t tma_stages() { static const uint32_t n = (uint32_t)env_int("GOOFY_B200_TMA_STAGES", 2, gb::kTmaMaxStages, 3); return n; }
uint32_t tma_rows() { static const uint32_t n = (uint32_t)env_int("GOOFY_B200_TMA_ROWS", 1, 4, 4); return n == 3u ? 4u : n; }

gb::FastDiv make_fastdiv(uint32_t d)
{
    gb::FastDiv f;
    f.d = d;
    f.m = ((1ull << 40) + d - 1u) / d;
    return f;
}

constexpr uint64_t kTmaMaxTiles = (1ull << 24) - 1u;   // FastDiv range

// Shapes the tile kernel's index arithmetic covers (FastDiv ranges, tensor-map limits).  Batches with more warp tiles
// than FastDiv covers are cut into several launches by launch_tma, so only one image has to fit.
bool tma_eligible(uint32_t bw, uint32_t bh, uint32_t stride, uint64_t srcPitch, uint32_t nImages)
{
    const uint64_t tilesX = (bw + gb::kTmaTileBlocksX - 1u) / gb::kTmaTileBlocksX;
    if (bh > 65536u || tilesX > 65536u || tilesX * bh > kTmaMaxTiles) return false;
    if (nImages > 1u && (srcPitch % 16u != 0u || srcPitch >= (1ull << 40))) return false;
    if ((uint64_t)stride * bh * 4u >= (1ull << 40)) return false;
    return tensor_map_encoder() != nullptr;
}

template <int MODE, int RB>
int launch_tma_r(void* dst, void* dst2, const void* src, uint32_t width, uint32_t height, uint32_t stride, uint64_t srcPitch,
                 uint64_t dstPitch, uint32_t nImages, cudaStream_t stream, int dev)
{
    constexpr int kLog = RB == 1 ? 0 : RB == 2 ? 1 : 2;
    const uint32_t nStages = tma_stages();
    const int smemBytes = (int)nStages * RB * gb::kTmaBlockRowBytes;
    int smCount = 0, ctasPerSm = 0;
    {
        std::lock_guard<std::mutex> g(g_devInfoMutex);
        DeviceInfo& di = g_devInfo[dev];
        if (di.smCount == 0) GB_CUDA(cudaDeviceGetAttribute(&di.smCount, cudaDevAttrMultiProcessorCount, dev));
        if (di.tmaCtasPerSm[MODE][kLog] == 0 || di.tmaSmemBytes[MODE][kLog] != smemBytes) {
            GB_CUDA(cudaFuncSetAttribute(gb::encode_tma_kernel<MODE, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
            int n = 0;
            GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gb::encode_tma_kernel<MODE, RB>, gb::tma_threads(RB), smemBytes));
            di.tmaCtasPerSm[MODE][kLog] = n > 0 ? n : 1;
            di.tmaSmemBytes[MODE][kLog] = smemBytes;
        }
        smCount = di.smCount;
        ctasPerSm = di.tmaCtasPerSm[MODE][kLog];
    }
    static const int promo = env_int("GOOFY_B200_TMA_L2PROMO", 0, 3, 3);
    static const uint32_t hint = (uint32_t)env_int("GOOFY_B200_TMA_HINT", 0, 1, 0);   // evict-first costs 4 % (round 1)
    // Persistent grid: tiles are dealt round-robin to the resident CTAs, whose rings keep them out of lock-step.
    // GOOFY_B200_TMA_GRID_MULT > 1 launches that many times more CTAs (each then walks fewer tiles).
    static const uint32_t gridMult = (uint32_t)env_int("GOOFY_B200_TMA_GRID_MULT", 1, 64, 1);

    const uint32_t bw = width / 4u, bh = height / 4u;
    const uint32_t tilesX = (bw + gb::kTmaTileBlocksX - 1u) / gb::kTmaTileBlocksX, groups = (bh + (uint32_t)RB - 1u) / (uint32_t)RB;
    const uint64_t tilesPerImage = (uint64_t)tilesX * groups;
    const uint32_t imagesPerLaunch = (uint32_t)std::min<uint64_t>(nImages, kTmaMaxTiles / tilesPerImage);
    for (uint32_t img0 = 0; img0 < nImages; img0 += imagesPerLaunch) {
        const uint32_t n = std::min(imagesPerLaunch, nImages - img0);
        const uint8_t* s0 = (const uint8_t*)src + (uint64_t)img0 * srcPitch;
        CUtensorMap map;
        const cuuint64_t dims[3] = {width, height, n};
        const cuuint64_t strides[2] = {stride, n > 1u ? srcPitch : (cuuint64_t)stride * height};
        const cuuint32_t box[3] = {(cuuint32_t)gb::kTmaBoxPixels, (cuuint32_t)(4 * RB), 1u};
        const cuuint32_t elemStrides[3] = {1u, 1u, 1u};
        const CUresult r = tensor_map_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(s0), dims, strides, box,
                                                elemStrides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return GOOFY_B200_E_ARGS;
        gb::TmaParams P;
        P.dst = (uint8_t*)dst + (uint64_t)img0 * dstPitch;
        P.dst2 = dst2 ? (uint8_t*)dst2 + (uint64_t)img0 * dstPitch : nullptr;
        P.dstPitch = dstPitch;
        P.bw = bw;
        P.bh = bh;
        P.nTiles = (uint32_t)(tilesPerImage * n);
        P.nStages = nStages;
        P.evictFirst = hint;
        P.tilesX = make_fastdiv(tilesX);
        P.groups = make_fastdiv(groups);
        uint32_t grid = (uint32_t)smCount * (uint32_t)ctasPerSm * gridMult;
        if (grid > P.nTiles) grid = P.nTiles;

        static const bool pdl = env_int("GOOFY_B200_PDL", 0, 1, 1) != 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid, 1, 1);
        cfg.blockDim = dim3(gb::tma_threads(RB), 1, 1);
        cfg.dynamicSmemBytes = (size_t)smemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 1 : 0;
        t_lastKernel = GB_KNAME("encode_tma_kernel");
        const cudaError_t e = cudaLaunchKernelEx(&cfg, gb::encode_tma_kernel<MODE, RB>, map, P);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    return GOOFY_B200_OK;
}

template <int MODE>
int launch_tma(void* dst, void* dst2, const void* src, uint32_t width, uint32_t height, uint32_t stride, uint64_t srcPitch,
               uint64_t dstPitch, uint32_t nImages, cudaStream_t stream, int dev)
{
    switch (tma_rows()) {
        case 1u: return launch_tma_r<MODE, 1>(dst, dst2, src, width, height, stride, srcPitch, dstPitch, nImages, stream, dev);
        case 2u: return launch_tma_r<MODE, 2>(dst, dst2, src, width, height, stride, srcPitch, dstPitch, nImages, stream, dev);
        default: return launch_tma_r<MODE, 4>(dst, dst2, src, width, height, stride, srcPitch, dstPitch, nImages, stream, dev);
    }
}

// Which load layer serves a uniform launch.  AUTO: see DESIGN.md section 3 ("load path policy").
bool choose_tma(uint32_t bw, uint32_t bh, uint32_t stride, uint64_t srcPitch, uint32_t nImages)
{
    const int path = g_loadPath.load(std::memory_order_relaxed);
    if (path == GOOFY_B200_LOAD_DIRECT) return false;
    if (!tma_eligible(bw, bh, stride, srcPitch, nImages)) return false;
    if (path == GOOFY_B200_LOAD_TMA) return true;
    return false;
}

// Images of a batch must not overlap: a pitch of 0 (or any pitch smaller than one image / one result) would make
// every image read the same pixels or overwrite the same blocks while the call still reports success.
bool pitches_cover_images(uint32_t width, uint32_t height, uint32_t stride, uint64_t srcPitch, uint64_t dstPitch)
{
    const uint64_t imageSpan = (uint64_t)(height - 1u) * stride + (uint64_t)width * 4u;
    return srcPitch >= imageSpan && dstPitch >= (uint64_t)width * height / 2u;
}

int encode_uniform(int mode, void* dst, void* dst2, const void* src, uint32_t width, uint32_t height, uint32_t stride,
                   uint64_t srcPitch, uint64_t dstPitch, uint32_t nImages, cudaStream_t stream)
{
    int rc = check_shape(width, height, stride);
    if (rc != GOOFY_B200_OK) return rc;
    if (width == 0u || height == 0u || nImages == 0u) return GOOFY_B200_OK;
    rc = check_pointers(src, dst);
    if (rc != GOOFY_B200_OK) return rc;
    if (mode == gb::kDual) {
        rc = check_pointers(src, dst2);
        if (rc != GOOFY_B200_OK) return rc;
    }
    if (nImages > 1u && ((srcPitch & 15u) != 0u || (dstPitch & 7u) != 0u)) return GOOFY_B200_E_ALIGN;
    if (nImages > 1u && !pitches_cover_images(width, height, stride, srcPitch, dstPitch)) return GOOFY_B200_E_ARGS;
    int dev = -1;
    rc = ensure_device_ready(&dev);
    if (rc != GOOFY_B200_OK) return rc;

    if (choose_tma(width / 4u, height / 4u, stride, srcPitch, nImages)) {
        switch (mode) {
            case gb::kDxt1: return launch_tma<gb::kDxt1>(dst, dst2, src, width, height, stride, srcPitch, dstPitch, nImages, stream, dev);
            case gb::kEtc1: return launch_tma<gb::kEtc1>(dst, dst2, src, width, height, stride, srcPitch, dstPitch, nImages, stream, dev);
            case gb::kDual: return launch_tma<gb::kDual>(dst, dst2, src, width, height, stride, srcPitch, dstPitch, nImages, stream, dev);
            default: return GOOFY_B200_E_CODEC;
        }
    }

    gb::EncodeParams P = {};
    P.src = (const uint8_t*)src;
    P.dst = (uint8_t*)dst;
    P.dst2 = (uint8_t*)dst2;
    P.bw = width / 4u;
    P.bh = height / 4u;
    P.stride = stride;
    P.by0 = 0;
    P.srcPitch = srcPitch;
    P.dstPitch = dstPitch;
    switch (mode) {
        case gb::kDxt1: return launch_direct<gb::kDxt1>(P, nImages, stream, dev);
        case gb::kEtc1: return launch_direct<gb::kEtc1>(P, nImages, stream, dev);
        case gb::kDual: return launch_direct<gb::kDual>(P, nImages, stream, dev);
        default: return GOOFY_B200_E_CODEC;
    }
}

// Float-reference flavour: one-shot CTAs; batches use grid.z (pitches are free-form).
int encode_floatref(int codec, void* dst, const void* src, uint32_t width, uint32_t height, uint32_t stride, uint64_t srcPitch,
                    uint64_t dstPitch, uint32_t nImages, cudaStream_t stream)
{
    int rc = check_shape_floatref(width, height, stride);
    if (rc != GOOFY_B200_OK) return rc;
    if (width == 0u || height == 0u || nImages == 0u) return GOOFY_B200_OK;
    rc = check_pointers(src, dst);
    if (rc != GOOFY_B200_OK) return rc;
    if (nImages > 1u && ((srcPitch & 15u) != 0u || (dstPitch & 7u) != 0u)) return GOOFY_B200_E_ALIGN;
    if (nImages > 1u && !pitches_cover_images(width, height, stride, srcPitch, dstPitch)) return GOOFY_B200_E_ARGS;
    rc = ensure_device_ready();
    if (rc != GOOFY_B200_OK) return rc;
    gb::EncodeParams P = {};
    P.src = (const uint8_t*)src;
    P.dst = (uint8_t*)dst;
    P.dst2 = nullptr;
    P.bw = width / 4u;
    P.bh = height / 4u;
    P.stride = stride;
    P.srcPitch = srcPitch;
    P.dstPitch = dstPitch;
    uint32_t tx = 32u;
    while (tx < 256u && tx < P.bw) tx <<= 1;
    const uint32_t ty = 256u / tx;
    const dim3 block(tx, ty, 1);
    const uint32_t gx = (P.bw + tx - 1u) / tx, rowGroups = (P.bh + ty - 1u) / ty;
    if (codec == GOOFY_B200_ETC1_FLOATREF) {
        // row-walking CTAs, four block rows each (the ETC1s flavour pays a table staging and a barrier per CTA)
        uint32_t gy = (rowGroups + 3u) / 4u;
        if (gy > 65535u) gy = 65535u;
        for (uint32_t img0 = 0; img0 < nImages; img0 += 65535u) {
            const uint32_t nz = nImages - img0 < 65535u ? nImages - img0 : 65535u;
            gb::EncodeParams Q = P;
            Q.by0 = 0;
            Q.src += (uint64_t)img0 * srcPitch;
            Q.dst += (uint64_t)img0 * dstPitch;
            t_lastKernel = "encode_floatref_kernel<etc1s, row-walking>";
            const int rc = launch_pdl(gb::encode_floatref_kernel<gb::kEtc1, true>, dim3(gx, gy, nz), block, stream, Q);
            if (rc != GOOFY_B200_OK) return rc;
        }
        return GOOFY_B200_OK;
    }
    const uint32_t rowsPerLaunch = 65535u * ty;
    for (uint32_t img0 = 0; img0 < nImages; img0 += 65535u) {
        const uint32_t nz = nImages - img0 < 65535u ? nImages - img0 : 65535u;
        for (uint32_t by0 = 0; by0 < P.bh; by0 += rowsPerLaunch) {
            const uint32_t rows = P.bh - by0 < rowsPerLaunch ? P.bh - by0 : rowsPerLaunch;
            gb::EncodeParams Q = P;
            Q.by0 = by0;
            Q.src += (uint64_t)img0 * srcPitch;
            Q.dst += (uint64_t)img0 * dstPitch;
            const dim3 grid(gx, (rows + ty - 1u) / ty, nz);
            t_lastKernel = "encode_floatref_kernel<dxt1>";
            const int rc = launch_pdl(gb::encode_floatref_kernel<gb::kDxt1, false>, grid, block, stream, Q);
            if (rc != GOOFY_B200_OK) return rc;
        }
    }
    return GOOFY_B200_OK;
}

// Packed RGB8 input (encode_kernels.cuh: encode_rgb24_kernel).  codec: DXT1, ETC1, BOTH (dst2 = the ETC1s blocks) or a
// float-reference flavour.  Rows and the base pointer only need 4-byte alignment (a tightly packed image with
// width % 4 == 0 qualifies); the shape rules are the codec's own.
int encode_rgb24(int codec, void* dst, void* dst2, const void* src, uint32_t width, uint32_t height, uint32_t stride, uint64_t srcPitch,
                 uint64_t dstPitch, uint32_t nImages, cudaStream_t stream)
{
    const bool both = codec == GOOFY_B200_BOTH;
    if (!both && !is_codec(codec)) return GOOFY_B200_E_CODEC;
    if (width % (is_floatref(codec) ? 4u : 16u) != 0u) return GOOFY_B200_E_WIDTH;
    if (height % 4u != 0u) return GOOFY_B200_E_HEIGHT;
    if (width == 0u || height == 0u || nImages == 0u) return GOOFY_B200_OK;
    if ((uint64_t)stride < (uint64_t)width * 3u) return GOOFY_B200_E_STRIDE;
    if (!src || !dst || (both && !dst2)) return GOOFY_B200_E_NULL;
    if (((uintptr_t)src & 3u) != 0u || (stride & 3u) != 0u || ((uintptr_t)dst & 7u) != 0u || (both && ((uintptr_t)dst2 & 7u) != 0u))
        return GOOFY_B200_E_ALIGN;
    if (nImages > 1u) {
        if ((srcPitch & 3u) != 0u || (dstPitch & 7u) != 0u) return GOOFY_B200_E_ALIGN;
        if (srcPitch < (uint64_t)(height - 1u) * stride + (uint64_t)width * 3u || dstPitch < (uint64_t)width * height / 2u) return GOOFY_B200_E_ARGS;
    }
    int dev = -1;
    int rc = ensure_device_ready(&dev);
    if (rc != GOOFY_B200_OK) return rc;
    const int sms = sm_count(dev);
    if (sms <= 0) return GOOFY_B200_E_DEVICE;

    gb::EncodeParams P = {};
    P.src = (const uint8_t*)src;
    P.dst = (uint8_t*)dst;
    P.dst2 = (uint8_t*)dst2;
    P.bw = width / 4u;
    P.bh = height / 4u;
    P.stride = stride;
    P.srcPitch = srcPitch;
    P.dstPitch = dstPitch;
    uint32_t tx = 32u;
    while (tx < 256u && tx < P.bw) tx <<= 1;
    const uint32_t ty = 256u / tx;
    const dim3 block(tx, ty, 1);
    const uint32_t gx = (P.bw + tx - 1u) / tx, rowGroups = (P.bh + ty - 1u) / ty;
    // a few block rows per CTA, never fewer CTAs than are resident at once (launch_rows has the measurements)
    const int mode = both ? gb::kDual : (codec == GOOFY_B200_ETC1 || codec == GOOFY_B200_ETC1_FLOATREF) ? gb::kEtc1 : gb::kDxt1;
    const uint32_t resident = (uint32_t)sms * (uint32_t)gb::rgb24_ctas_per_sm(mode);
    static const uint32_t rowsEnv = (uint32_t)env_int("GOOFY_B200_RGB24_ROWS_PER_CTA", 1, 1 << 20, 0);
    // block rows per CTA (session V, 4 x 8192^2): DXT1 6449 / 6719 / 6817 / 6843 / 6799 GB/s at 2 / 3 / 4 / 6 / 8, ETC1s 5362 / 5551 /
    // 5646 / 5736 / 5727, both 5233 / 5324 / 5418 / 5456 / 5566
    const uint32_t rowsPerCta = rowsEnv ? rowsEnv : mode == gb::kDual ? 8u : mode == gb::kEtc1 ? 6u : 4u;
    uint32_t gy = (rowGroups + rowsPerCta - 1u) / rowsPerCta;
    const uint32_t perImage = (resident + nImages - 1u) / nImages;
    if (gy < perImage / gx) gy = perImage / gx;
    if (gy == 0u) gy = 1u;
    if (gy > rowGroups) gy = rowGroups;
    if (gy > 65535u) gy = 65535u;
    for (uint32_t img0 = 0; img0 < nImages; img0 += 65535u) {
        const uint32_t nz = nImages - img0 < 65535u ? nImages - img0 : 65535u;
        gb::EncodeParams Q = P;
        Q.src += (uint64_t)img0 * srcPitch;
        Q.dst += (uint64_t)img0 * dstPitch;
        if (Q.dst2) Q.dst2 += (uint64_t)img0 * dstPitch;
        const dim3 grid(gx, gy, nz);
        switch (codec) {
            case GOOFY_B200_DXT1: rc = launch_encode(gb::encode_rgb24_kernel<gb::kDxt1, 0>, grid, block, stream, Q, "encode_rgb24_kernel<dxt1>"); break;
            case GOOFY_B200_ETC1: rc = launch_encode(gb::encode_rgb24_kernel<gb::kEtc1, 0>, grid, block, stream, Q, "encode_rgb24_kernel<etc1s>"); break;
            case GOOFY_B200_BOTH: rc = launch_encode(gb::encode_rgb24_kernel<gb::kDual, 0>, grid, block, stream, Q, "encode_rgb24_kernel<dxt1+etc1s>"); break;
            case GOOFY_B200_DXT1_FLOATREF: rc = launch_encode(gb::encode_rgb24_kernel<gb::kDxt1, 1>, grid, block, stream, Q, "encode_rgb24_kernel<dxt1, float reference>"); break;
            default: rc = launch_encode(gb::encode_rgb24_kernel<gb::kEtc1, 1>, grid, block, stream, Q, "encode_rgb24_kernel<etc1s, float reference>"); break;
        }
        if (rc != GOOFY_B200_OK) return rc;
    }
    return GOOFY_B200_OK;
}

// Device-resident dispatch by codec selector (SSE2-exact or float-reference-exact flavour).
int encode_any(int codec, void* dst, const void* src, uint32_t width, uint32_t height, uint32_t stride, uint64_t srcPitch,
               uint64_t dstPitch, uint32_t nImages, cudaStream_t stream)
{
    if (!is_codec(codec)) return GOOFY_B200_E_CODEC;
    if (is_floatref(codec)) return encode_floatref(codec, dst, src, width, height, stride, srcPitch, dstPitch, nImages, stream);
    return encode_uniform(codec, dst, nullptr, src, width, height, stride, srcPitch, dstPitch, nImages, stream);
}

}  // namespace
