// host_launch.cuh -- launch configuration of the device-resident entry points: load-layer policy (one-shot,
// row-walking, cp.async ring, TMA tiles), programmatic dependent launch, the float-reference flavour.
#pragma once
#include "host_common.cuh"

namespace {

template <typename Kernel>
int launch_encode(Kernel kernel, dim3 grid, dim3 block, cudaStream_t stream, const gb::EncodeParams& P)
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
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, P);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(e);
}

template <int MODE, bool PITCHED>
int launch_direct_grid(const gb::EncodeParams& Q, dim3 grid, dim3 block, cudaStream_t stream)
{
    // 32-bit in-image offsets unless the image spans 4 GiB or more
    if ((uint64_t)Q.bh * 4u * Q.stride + (uint64_t)Q.bw * 16u < 0xFFFFFFFFull)
        return launch_encode(gb::encode_direct_kernel<MODE, false, PITCHED>, grid, block, stream, Q);
    return launch_encode(gb::encode_direct_kernel<MODE, true, PITCHED>, grid, block, stream, Q);
}

int sm_count(int dev);

// Persistent row-walking launch for one (possibly very tall) image.
template <int MODE>
int launch_rows(const gb::EncodeParams& P, cudaStream_t stream, int dev)
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
    const bool async = loadPath == GOOFY_B200_LOAD_ASYNC || (loadPath == GOOFY_B200_LOAD_AUTO && MODE == gb::kDual);
    const uint32_t resident = async ? (uint32_t)sms * (uint32_t)GB_ASYNC_CTAS(MODE)
                                    : (uint32_t)sms * (uint32_t)gb::ctas_per_sm(MODE) * (256u / (uint32_t)GB_TPB);
    // Each CTA walks a few block rows: enough to amortise the per-thread set-up, few enough that CTAs keep
    // retiring and restarting at staggered times (fully persistent CTAs run in lock-step and are 10 % slower;
    // profiles/r01_rows_grid_sweep.txt).  Never fewer CTAs than one resident wave.
    // Measured on the batched 4 x 8192^2 launch: ETC1s 6915 / 7011 / 7010 / 6966 / 6892 GB/s at 2 / 3 / 4 / 6 / 8 rows
    // per CTA; dual-output 5835 / 5932 / 6055 / 6193 / 6293.
    static const uint32_t rowsEnv = []() { const char* e = getenv("GOOFY_B200_ROWS_PER_CTA"); const int v = e ? atoi(e) : 0; return v > 0 ? (uint32_t)v : 0u; }();
    const uint32_t rowsPerCta = rowsEnv ? rowsEnv : (MODE == gb::kDual ? 8u : 4u);
    uint32_t gy = (rowGroups + rowsPerCta - 1u) / rowsPerCta;
    if (gy < resident / gx) gy = resident / gx;
    if (gy == 0u) gy = 1u;
    if (gy > rowGroups) gy = rowGroups;
    if (gy > 65535u) gy = 65535u;
    const dim3 grid(gx, gy, 1), block(tx, ty, 1);
    const bool narrow = (uint64_t)P.bh * 4u * P.stride + (uint64_t)P.bw * 16u < 0xFFFFFFFFull;
    if (async)
        return narrow ? launch_encode(gb::encode_rows_async_kernel<MODE, false>, grid, block, stream, P)
                      : launch_encode(gb::encode_rows_async_kernel<MODE, true>, grid, block, stream, P);
    if (narrow) return launch_encode(gb::encode_rows_kernel<MODE, false>, grid, block, stream, P);
    return launch_encode(gb::encode_rows_kernel<MODE, true>, grid, block, stream, P);
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
    // gain 4-13 % from row-walking CTAs (dual-output: through the cp.async ring, see launch_rows).
    const int path = g_loadPath.load(std::memory_order_relaxed);
    const bool rows = path == GOOFY_B200_LOAD_DIRECT || path == GOOFY_B200_LOAD_ASYNC ||
                      (path == GOOFY_B200_LOAD_AUTO && MODE != gb::kDxt1);
    if (nImages == 1u && rows) return launch_rows<MODE>(P, stream, dev);
    // one-shot CTAs (pitched batches): threads: x walks blocks along a row (coalescing), y stacks block rows for narrow images
    // (x is a power of two and x*y == GB_TPB: the kernels rely on exactly GB_TPB threads)
    uint32_t tx = 32u;
    while (tx < (uint32_t)GB_TPB && tx < P.bw) tx <<= 1;
    const uint32_t ty = (uint32_t)GB_TPB / tx;
    const dim3 block(tx, ty, 1);
    const uint32_t gx = (P.bw + tx - 1u) / tx;
    const uint32_t rowsPerLaunch = 65535u * ty;
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
    int tmaCtasPerSm[3] = {0, 0, 0};  // 0 = not yet configured
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

// Depth of the tile ring.  GOOFY_B200_TMA_STAGES overrides it for experiments (2..8).
uint32_t tma_stages()
{
    static const uint32_t n = []() -> uint32_t {
        const char* e = getenv("GOOFY_B200_TMA_STAGES");
        const int v = e ? atoi(e) : 0;
        return (v >= 2 && v <= gb::kTmaMaxStages) ? (uint32_t)v : 2u;
    }();
    return n;
}
gb::FastDiv make_fastdiv(uint32_t d)
{
    gb::FastDiv f;
    f.d = d;
    f.m = ((1ull << 40) + d - 1u) / d;
    return f;
}

// Shapes the tile kernel's index arithmetic covers (FastDiv ranges, tensor-map limits).
bool tma_eligible(uint32_t bw, uint32_t bh, uint32_t stride, uint64_t srcPitch, uint32_t nImages)
{
    const uint64_t tilesX = (bw + gb::kTmaThreads - 1u) / gb::kTmaThreads;
    const uint64_t nTiles = tilesX * bh * nImages;
    if (bh > 65536u || tilesX > 65536u || nTiles >= (1ull << 24)) return false;
    if (nImages > 1u && (srcPitch % 16u != 0u || srcPitch >= (1ull << 40))) return false;
    if ((uint64_t)stride * bh * 4u >= (1ull << 40)) return false;
    return tensor_map_encoder() != nullptr;
}

template <int MODE>
int launch_tma(void* dst, void* dst2, const void* src, uint32_t width, uint32_t height, uint32_t stride, uint64_t srcPitch,
               uint64_t dstPitch, uint32_t nImages, cudaStream_t stream, int dev)
{
    CUtensorMap map;
    const cuuint64_t dims[3] = {width, height, nImages};
    const cuuint64_t strides[2] = {stride, nImages > 1u ? srcPitch : (cuuint64_t)stride * height};
    const cuuint32_t box[3] = {(cuuint32_t)gb::kTmaBoxPixels, 4u, 1u};
    const cuuint32_t elemStrides[3] = {1u, 1u, 1u};
    static const int promo = []() { const char* e = getenv("GOOFY_B200_TMA_L2PROMO"); const int v = e ? atoi(e) : 3; return (v >= 0 && v <= 3) ? v : 3; }();
    const CUresult r = tensor_map_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(src), dims, strides, box,
                                            elemStrides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return GOOFY_B200_E_ARGS;

    const uint32_t nStages = tma_stages();
    const int smemBytes = (int)nStages * gb::kTmaStageBytes;
    int smCount = 0, ctasPerSm = 0;
    {
        std::lock_guard<std::mutex> g(g_devInfoMutex);
        DeviceInfo& di = g_devInfo[dev];
        if (di.smCount == 0) GB_CUDA(cudaDeviceGetAttribute(&di.smCount, cudaDevAttrMultiProcessorCount, dev));
        if (di.tmaCtasPerSm[MODE] == 0) {
            GB_CUDA(cudaFuncSetAttribute(gb::encode_tma_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
            int n = 0;
            GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gb::encode_tma_kernel<MODE>, gb::kTmaThreads, smemBytes));
            di.tmaCtasPerSm[MODE] = n > 0 ? n : 1;
        }
        smCount = di.smCount;
        ctasPerSm = di.tmaCtasPerSm[MODE];
    }
    gb::TmaParams P;
    P.dst = (uint8_t*)dst;
    P.dst2 = (uint8_t*)dst2;
    P.dstPitch = dstPitch;
    P.bw = width / 4u;
    P.bh = height / 4u;
    const uint32_t tilesX = (P.bw + gb::kTmaThreads - 1u) / gb::kTmaThreads;
    P.nTiles = tilesX * P.bh * nImages;
    P.nStages = nStages;
    static const uint32_t hint = []() { const char* e = getenv("GOOFY_B200_TMA_HINT"); return (e && e[0] == '1') ? 1u : 0u; }();
    P.evictFirst = hint;  // off by default: the evict-first policy costs 4 % (6543 vs 6815 GB/s, DXT1)
    P.tilesX = make_fastdiv(tilesX);
    P.rows = make_fastdiv(P.bh);
    // CTAs walk a few tiles each: a multiple of what is resident at once (fully persistent CTAs run in
    // lock-step and are slower, as with the row-walking kernels), never more than there are tiles
    static const uint32_t gridMult = []() { const char* e = getenv("GOOFY_B200_TMA_GRID_MULT"); const int v = e ? atoi(e) : 8; return v > 0 ? (uint32_t)v : 8u; }();
    uint32_t grid = (uint32_t)smCount * (uint32_t)ctasPerSm * gridMult;
    if (grid > P.nTiles) grid = P.nTiles;
    gb::encode_tma_kernel<MODE><<<grid, gb::kTmaThreads, smemBytes, stream>>>(map, P);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(cudaGetLastError());
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

    gb::EncodeParams P;
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
    rc = ensure_device_ready();
    if (rc != GOOFY_B200_OK) return rc;
    gb::EncodeParams P;
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
    const uint32_t gx = (P.bw + tx - 1u) / tx, rowsPerLaunch = 65535u * ty;
    for (uint32_t img0 = 0; img0 < nImages; img0 += 65535u) {
        const uint32_t nz = nImages - img0 < 65535u ? nImages - img0 : 65535u;
        for (uint32_t by0 = 0; by0 < P.bh; by0 += rowsPerLaunch) {
            const uint32_t rows = P.bh - by0 < rowsPerLaunch ? P.bh - by0 : rowsPerLaunch;
            gb::EncodeParams Q = P;
            Q.by0 = by0;
            Q.src += (uint64_t)img0 * srcPitch;
            Q.dst += (uint64_t)img0 * dstPitch;
            const dim3 grid(gx, (rows + ty - 1u) / ty, nz);
            if (codec == GOOFY_B200_DXT1_FLOATREF) gb::encode_floatref_kernel<gb::kDxt1><<<grid, block, 0, stream>>>(Q);
            else gb::encode_floatref_kernel<gb::kEtc1><<<grid, block, 0, stream>>>(Q);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            GB_CUDA(cudaGetLastError());
        }
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
