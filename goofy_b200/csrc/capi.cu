// capi.cu -- the extern "C" layer of libgoofy_b200.so (declared in include/goofy_b200.h).
//
// Thin by design: argument checks that mirror goofy::compressDXT1/ETC1
// (GoofyTC/goofy_tc.h:1497-1557), launch configuration, the host-pointer staging pipeline and
// the multi-GPU shard scheduler.  All arithmetic lives in block_codec.cuh.  There is no CPU
// encoder anywhere in this library: without a CUDA device every call returns an error code.
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/goofy_b200.h"
#include "encode_kernels.cuh"
#include "tma_kernels.cuh"
#include "decode_kernels.cuh"

namespace {

std::atomic<uint64_t> g_launches{0};

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? GOOFY_B200_OK : GOOFY_B200_E_CUDA_BASE - (int)e; }

#define GB_CUDA(call)                                   \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return cuda_rc(e__);    \
    } while (0)

std::atomic<int> g_loadPath{GOOFY_B200_LOAD_AUTO};

// ---- per-device one-time state: the ETC1 control table in device memory ----
constexpr int kMaxDevices = 64;
std::once_flag g_lutOnce[kMaxDevices];
int g_lutStatus[kMaxDevices];

int ensure_device_ready(int* deviceOut = nullptr)
{
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cudaGetLastError(); return GOOFY_B200_E_DEVICE; }
    if (dev < 0 || dev >= kMaxDevices) return GOOFY_B200_E_DEVICE;
    std::call_once(g_lutOnce[dev], [dev]() {
        gb::fill_control_lut_kernel<<<1, 256>>>();
        gb::fill_control_lut_ref_kernel<<<1, 256>>>();
        cudaError_t le = cudaGetLastError();
        if (le == cudaSuccess) le = cudaDeviceSynchronize();
        g_lutStatus[dev] = cuda_rc(le);
    });
    if (deviceOut) *deviceOut = dev;
    return g_lutStatus[dev];
}

// Shape checks in the reference's order (goofy_tc.h:1500-1508), then the new ones.
int check_shape(uint32_t width, uint32_t height, uint32_t stride)
{
    if (width % 16u != 0u) return GOOFY_B200_E_WIDTH;
    if (height % 4u != 0u) return GOOFY_B200_E_HEIGHT;
    if (width == 0u || height == 0u) return GOOFY_B200_OK;
    if ((uint64_t)stride < (uint64_t)width * 4u) return GOOFY_B200_E_STRIDE;
    if (stride % 16u != 0u) return GOOFY_B200_E_ALIGN;
    return GOOFY_B200_OK;
}

bool is_floatref(int codec) { return codec == GOOFY_B200_DXT1_FLOATREF || codec == GOOFY_B200_ETC1_FLOATREF; }
bool is_codec(int codec) { return codec == GOOFY_B200_DXT1 || codec == GOOFY_B200_ETC1 || is_floatref(codec); }

// goofyRef:: accepts any width that is a multiple of 4 (Src/goofy_tc_reference.cpp:796-804)
int check_shape_floatref(uint32_t width, uint32_t height, uint32_t stride)
{
    if (width % 4u != 0u) return GOOFY_B200_E_WIDTH;
    if (height % 4u != 0u) return GOOFY_B200_E_HEIGHT;
    if (width == 0u || height == 0u) return GOOFY_B200_OK;
    if ((uint64_t)stride < (uint64_t)width * 4u) return GOOFY_B200_E_STRIDE;
    if (stride % 16u != 0u) return GOOFY_B200_E_ALIGN;
    return GOOFY_B200_OK;
}

int check_pointers(const void* src, const void* dst)
{
    if (!src || !dst) return GOOFY_B200_E_NULL;
    if (((uintptr_t)src & 15u) != 0u || ((uintptr_t)dst & 7u) != 0u) return GOOFY_B200_E_ALIGN;
    return GOOFY_B200_OK;
}

// Launch with programmatic stream serialisation (see pdl_wait in encode_kernels.cuh): the kernel's CTAs
// may be scheduled while the previous kernel of the stream drains, then wait for it before touching memory.
// GOOFY_B200_PDL=0 turns it off.
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
    const bool async = g_loadPath.load(std::memory_order_relaxed) == GOOFY_B200_LOAD_ASYNC;
    const uint32_t resident = async ? (uint32_t)sms * (MODE == gb::kDual ? 5u : 6u)
                                    : (uint32_t)sms * (MODE == gb::kDual ? (uint32_t)GB_DUAL_CTAS : 8u) * (256u / (uint32_t)GB_TPB);
    // CTAs walk ~3.5 block rows each on an 8192^2 texture: enough to amortise the per-thread set-up,
    // few enough that CTAs keep retiring and restarting at staggered times (measured: 1x resident
    // 5634, 4x 6111, 14x 5640 GB/s for ETC1s; profiles/r01_rows_grid_sweep.txt).
    static const uint32_t gyMult = []() { const char* e = getenv("GOOFY_B200_ROWS_GY_MULT"); const int v = e ? atoi(e) : 4; return v > 0 ? (uint32_t)v : 4u; }();
    uint32_t gy = (uint32_t)(((uint64_t)resident * gyMult) / gx);
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
    // one-shot CTAs are marginally faster (6617 vs 6598 GB/s); the ETC1s and dual-output kernels are
    // instruction-bound and gain 4-13 % from row-walking CTAs.
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

// ---------------------------------------------------------------- host-pointer pipeline
// The image is cut into strips of whole block rows; strip i runs H2D -> kernel -> D2H on
// stream i % kSlots so the copies of neighbouring strips overlap each other and the kernels.
// Device scratch is cached per host thread and device and only ever grows.
constexpr int kSlots = 3;
constexpr size_t kStripBytes = 16u << 20;

struct HostPipe {
    int device = -1;
    cudaStream_t stream[kSlots] = {};
    void* dIn[kSlots] = {};
    void* dOut[kSlots] = {};
    size_t capIn = 0, capOut = 0;
    bool ready = false;

    int prepare(int dev, size_t needIn, size_t needOut)
    {
        if (ready && dev != device) release();
        if (!ready) {
            for (int i = 0; i < kSlots; ++i) GB_CUDA(cudaStreamCreateWithFlags(&stream[i], cudaStreamNonBlocking));
            device = dev;
            ready = true;
        }
        if (needIn > capIn) {
            capIn = 0;  // stays 0 if an allocation below fails, so the next call starts over
            for (int i = 0; i < kSlots; ++i) {
                if (dIn[i]) cudaFree(dIn[i]);
                dIn[i] = nullptr;
                GB_CUDA(cudaMalloc(&dIn[i], needIn));
            }
            capIn = needIn;
        }
        if (needOut > capOut) {
            capOut = 0;
            for (int i = 0; i < kSlots; ++i) {
                if (dOut[i]) cudaFree(dOut[i]);
                dOut[i] = nullptr;
                GB_CUDA(cudaMalloc(&dOut[i], needOut));
            }
            capOut = needOut;
        }
        return GOOFY_B200_OK;
    }
    void release()
    {
        for (int i = 0; i < kSlots; ++i) {
            if (dIn[i]) cudaFree(dIn[i]);
            if (dOut[i]) cudaFree(dOut[i]);
            if (stream[i]) cudaStreamDestroy(stream[i]);
            dIn[i] = dOut[i] = nullptr;
            stream[i] = nullptr;
        }
        capIn = capOut = 0;
        ready = false;
    }
    // No destructor on purpose: thread_local teardown can run after the CUDA runtime has
    // shut down; the driver reclaims everything at process exit.
};

thread_local HostPipe t_pipe;

// Pageable (malloc'd) host buffers cannot be DMA'd directly; the CUDA driver then stages them through
// one small internal buffer at ~10 GB/s.  The library stages them itself instead: a few persistent host
// threads copy each strip into pinned memory in parallel while the previous strips are in flight.
class CopyPool {
public:
    static CopyPool& get()
    {
        static CopyPool* pool = new CopyPool();  // leaked on purpose (see HostPipe)
        return *pool;
    }
    // dst/src rows of `rowBytes`, `rows` of them; the calling thread takes a share of the rows too
    void copy2d(uint8_t* dst, size_t dstPitch, const uint8_t* src, size_t srcPitch, size_t rowBytes, size_t rows)
    {
        if (rows * rowBytes < (1u << 20) || workers_.empty()) {
            run(dst, dstPitch, src, srcPitch, rowBytes, 0, rows);
            return;
        }
        std::lock_guard<std::mutex> serial(jobMutex_);  // one copy job at a time
        {
            std::lock_guard<std::mutex> g(m_);
            dst_ = dst; dstPitch_ = dstPitch; src_ = src; srcPitch_ = srcPitch; rowBytes_ = rowBytes; rows_ = rows;
            pending_ = (int)workers_.size();
            ++generation_;
        }
        cv_.notify_all();
        const size_t parts = workers_.size() + 1;
        run(dst, dstPitch, src, srcPitch, rowBytes, rows * (parts - 1) / parts, rows);  // the caller's share: the last slice
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
    }

    void copy1d(uint8_t* dst, const uint8_t* src, size_t bytes)
    {
        const size_t chunk = 1u << 16, full = bytes / chunk;
        if (full) copy2d(dst, chunk, src, chunk, chunk, full);
        if (bytes > full * chunk) std::memcpy(dst + full * chunk, src + full * chunk, bytes - full * chunk);
    }

private:
    CopyPool()
    {
        unsigned n = std::thread::hardware_concurrency();
        n = n > 16u ? 7u : (n > 2u ? n / 2u - 1u : 0u);  // plus the calling thread
        for (unsigned i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i); });
        for (auto& t : workers_) t.detach();
    }
    static void run(uint8_t* dst, size_t dstPitch, const uint8_t* src, size_t srcPitch, size_t rowBytes, size_t r0, size_t r1)
    {
        if (dstPitch == rowBytes && srcPitch == rowBytes) {
            std::memcpy(dst + r0 * rowBytes, src + r0 * rowBytes, (r1 - r0) * rowBytes);
            return;
        }
        for (size_t r = r0; r < r1; ++r) std::memcpy(dst + r * dstPitch, src + r * srcPitch, rowBytes);
    }
    void loop(unsigned index)
    {
        uint64_t seen = 0;
        for (;;) {
            uint8_t* dst; const uint8_t* src; size_t dp, sp, rb, rows;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return generation_ != seen; });
                seen = generation_;
                dst = dst_; src = src_; dp = dstPitch_; sp = srcPitch_; rb = rowBytes_; rows = rows_;
            }
            const size_t parts = workers_.size() + 1;
            run(dst, dp, src, sp, rb, rows * index / parts, rows * (index + 1) / parts);
            {
                std::lock_guard<std::mutex> g(m_);
                --pending_;
            }
            done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex jobMutex_, m_;
    std::condition_variable cv_, done_;
    uint8_t* dst_ = nullptr;
    const uint8_t* src_ = nullptr;
    size_t dstPitch_ = 0, srcPitch_ = 0, rowBytes_ = 0, rows_ = 0;
    int pending_ = 0;
    uint64_t generation_ = 0;
};

// Pinned staging strips, allocated only when a pageable buffer is first seen by this thread.
struct HostStage {
    void* in[kSlots] = {};
    void* out[kSlots] = {};
    size_t capIn = 0, capOut = 0;
    int ensure(size_t needIn, size_t needOut)
    {
        if (needIn > capIn) {
            capIn = 0;
            for (int i = 0; i < kSlots; ++i) {
                if (in[i]) cudaFreeHost(in[i]);
                in[i] = nullptr;
                GB_CUDA(cudaHostAlloc(&in[i], needIn, cudaHostAllocDefault));
            }
            capIn = needIn;
        }
        if (needOut > capOut) {
            capOut = 0;
            for (int i = 0; i < kSlots; ++i) {
                if (out[i]) cudaFreeHost(out[i]);
                out[i] = nullptr;
                GB_CUDA(cudaHostAlloc(&out[i], needOut, cudaHostAllocDefault));
            }
            capOut = needOut;
        }
        return GOOFY_B200_OK;
    }
};
thread_local HostStage t_stage;

bool is_pageable(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

int encode_host(int codec, void* result, const void* input, uint32_t width, uint32_t height, uint32_t stride)
{
    if (!is_codec(codec)) return GOOFY_B200_E_CODEC;
    int rc = is_floatref(codec) ? check_shape_floatref(width, height, stride) : check_shape(width, height, stride);
    if (rc != GOOFY_B200_OK) return rc;
    if (width == 0u || height == 0u) return GOOFY_B200_OK;
    if (!result || !input) return GOOFY_B200_E_NULL;
    if (((uintptr_t)input & 15u) != 0u) return GOOFY_B200_E_ALIGN;  // the reference's aligned-load contract
    int dev = -1;
    rc = ensure_device_ready(&dev);
    if (rc != GOOFY_B200_OK) return rc;

    const size_t rowBytes = (size_t)width * 4u;
    const uint32_t blockRows = height / 4u;
    uint32_t stripRows = (uint32_t)(kStripBytes / (rowBytes * 4u));
    if (stripRows == 0u) stripRows = 1u;
    if (stripRows > blockRows) stripRows = blockRows;
    const size_t outRowBytes = (size_t)(width / 4u) * 8u;
    const size_t stripIn = (size_t)stripRows * 4u * rowBytes, stripOut = (size_t)stripRows * outRowBytes;
    rc = t_pipe.prepare(dev, stripIn, stripOut);
    if (rc != GOOFY_B200_OK) return rc;

    // Pinned buffers are DMA'd in place; pageable ones go through the pinned staging strips.
    const bool stageIn = is_pageable(input), stageOut = is_pageable(result);
    if (stageIn || stageOut) {
        rc = t_stage.ensure(stageIn ? stripIn : 0, stageOut ? stripOut : 0);
        if (rc != GOOFY_B200_OK) return rc;
    }
    struct Pending { uint32_t r0 = 0, rows = 0; bool live = false; } pending[kSlots];
    auto retire = [&](int slot) -> int {  // wait for the slot's strip and hand its blocks to the caller
        if (!pending[slot].live) return GOOFY_B200_OK;
        GB_CUDA(cudaStreamSynchronize(t_pipe.stream[slot]));
        if (stageOut)
            CopyPool::get().copy1d((uint8_t*)result + (size_t)pending[slot].r0 * outRowBytes, (const uint8_t*)t_stage.out[slot],
                                   (size_t)pending[slot].rows * outRowBytes);
        pending[slot].live = false;
        return GOOFY_B200_OK;
    };

    int slot = 0;
    for (uint32_t r0 = 0; r0 < blockRows; r0 += stripRows, slot = (slot + 1) % kSlots) {
        const uint32_t rows = blockRows - r0 < stripRows ? blockRows - r0 : stripRows;
        cudaStream_t s = t_pipe.stream[slot];
        const uint8_t* src = (const uint8_t*)input + (size_t)r0 * 4u * stride;
        if (stageIn || stageOut) {
            rc = retire(slot);  // the staging strips of this slot are about to be reused
            if (rc != GOOFY_B200_OK) return rc;
        }
        if (stageIn) {
            CopyPool::get().copy2d((uint8_t*)t_stage.in[slot], rowBytes, src, stride, rowBytes, (size_t)rows * 4u);
            GB_CUDA(cudaMemcpyAsync(t_pipe.dIn[slot], t_stage.in[slot], (size_t)rows * 4u * rowBytes, cudaMemcpyHostToDevice, s));
        } else {
            // stream order protects the slot's device scratch: its previous strip finished D2H on the same stream
            GB_CUDA(cudaMemcpy2DAsync(t_pipe.dIn[slot], rowBytes, src, stride, rowBytes, (size_t)rows * 4u, cudaMemcpyHostToDevice, s));
        }
        rc = encode_any(codec, t_pipe.dOut[slot], t_pipe.dIn[slot], width, rows * 4u, (uint32_t)rowBytes, 0, 0, 1, s);
        if (rc != GOOFY_B200_OK) return rc;
        GB_CUDA(cudaMemcpyAsync(stageOut ? t_stage.out[slot] : (void*)((uint8_t*)result + (size_t)r0 * outRowBytes), t_pipe.dOut[slot],
                                (size_t)rows * outRowBytes, cudaMemcpyDeviceToHost, s));
        pending[slot].r0 = r0;
        pending[slot].rows = rows;
        pending[slot].live = true;
    }
    // drain in strip order (the oldest outstanding strip is in the slot the loop would use next)
    for (int i = 0; i < kSlots; ++i, slot = (slot + 1) % kSlots) {
        rc = retire(slot);
        if (rc != GOOFY_B200_OK) return rc;
    }
    return GOOFY_B200_OK;
}

// ---------------------------------------------------------------- ragged batch
template <int MODE>
int launch_batch(const gb::BatchImage* dImages, const uint32_t* dStart, uint32_t n, uint32_t totalCtas, cudaStream_t stream)
{
    gb::encode_batch_kernel<MODE><<<totalCtas, dim3(gb::kBatchTileX, gb::kBatchTileY, 1), 0, stream>>>(dImages, dStart, n);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(cudaGetLastError());
}

// Descriptor tables live in a small per-thread device arena that is recycled in stream order.
struct BatchArena {
    void* dev = nullptr;
    void* host = nullptr;  // pinned
    size_t cap = 0;
    int device = -1;
    cudaEvent_t done = nullptr;
};
thread_local BatchArena t_arena;

int encode_batch_current_device(int codec, const GoofyB200Image* descs, const uint32_t* order, uint32_t n, cudaStream_t stream)
{
    if (codec != GOOFY_B200_DXT1 && codec != GOOFY_B200_ETC1) return GOOFY_B200_E_CODEC;
    if (n == 0u) return GOOFY_B200_OK;
    if (!descs) return GOOFY_B200_E_NULL;
    int dev = -1;
    int rc = ensure_device_ready(&dev);
    if (rc != GOOFY_B200_OK) return rc;

    std::vector<gb::BatchImage> images;
    std::vector<uint32_t> start;
    images.reserve(n);
    start.reserve(n + 1);
    uint64_t total = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const GoofyB200Image& d = descs[order ? order[k] : k];
        rc = check_shape(d.width, d.height, d.stride);
        if (rc != GOOFY_B200_OK) return rc;
        if (d.width == 0u || d.height == 0u) continue;
        rc = check_pointers(d.src, d.dst);
        if (rc != GOOFY_B200_OK) return rc;
        gb::BatchImage im;
        im.src = (const uint8_t*)d.src;
        im.dst = (uint8_t*)d.dst;
        im.bw = d.width / 4u;
        im.bh = d.height / 4u;
        im.stride = d.stride;
        im.tilesX = (im.bw + gb::kBatchTileX - 1u) / gb::kBatchTileX;
        start.push_back((uint32_t)total);
        total += (uint64_t)im.tilesX * ((im.bh + gb::kBatchTileY - 1u) / gb::kBatchTileY);
        if (total > 0x7FFFFFFFull) return GOOFY_B200_E_ARGS;
        images.push_back(im);
    }
    if (images.empty()) return GOOFY_B200_OK;
    const uint32_t m = (uint32_t)images.size();
    const size_t bytesImages = (size_t)m * sizeof(gb::BatchImage);
    const size_t bytes = bytesImages + (size_t)m * sizeof(uint32_t);

    BatchArena& A = t_arena;
    if (A.device != dev || bytes > A.cap) {
        if (A.done) { cudaEventSynchronize(A.done); }
        if (A.dev) cudaFree(A.dev);
        if (A.host) cudaFreeHost(A.host);
        A.dev = A.host = nullptr;
        A.cap = 0;
        size_t cap = bytes < (1u << 16) ? (1u << 16) : bytes * 2u;
        GB_CUDA(cudaMalloc(&A.dev, cap));
        GB_CUDA(cudaHostAlloc(&A.host, cap, cudaHostAllocDefault));
        if (!A.done) GB_CUDA(cudaEventCreateWithFlags(&A.done, cudaEventDisableTiming));
        A.cap = cap;
        A.device = dev;
    } else if (A.done) {
        GB_CUDA(cudaEventSynchronize(A.done));  // previous batch has consumed the table
    }
    std::memcpy(A.host, images.data(), bytesImages);
    std::memcpy((uint8_t*)A.host + bytesImages, start.data(), (size_t)m * sizeof(uint32_t));
    GB_CUDA(cudaMemcpyAsync(A.dev, A.host, bytes, cudaMemcpyHostToDevice, stream));
    const gb::BatchImage* dImages = (const gb::BatchImage*)A.dev;
    const uint32_t* dStart = (const uint32_t*)((const uint8_t*)A.dev + bytesImages);
    rc = codec == GOOFY_B200_DXT1 ? launch_batch<gb::kDxt1>(dImages, dStart, m, (uint32_t)total, stream)
                                  : launch_batch<gb::kEtc1>(dImages, dStart, m, (uint32_t)total, stream);
    if (rc != GOOFY_B200_OK) return rc;
    GB_CUDA(cudaEventRecord(A.done, stream));
    return GOOFY_B200_OK;
}

// ---------------------------------------------------------------- shard scheduler
// One persistent host thread per device.  A job is a closure run with that device current;
// there is no cross-device communication of any kind (blocks are independent).
class DeviceWorker {
public:
    explicit DeviceWorker(int device) : device_(device), thread_([this] { loop(); }) {}
    ~DeviceWorker()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        thread_.join();
    }
    void submit(std::function<int()> job)
    {
        {
            std::lock_guard<std::mutex> g(m_);
            job_ = std::move(job);
            hasJob_ = true;
            done_ = false;
        }
        cv_.notify_all();
    }
    int wait()
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this] { return done_; });
        return rc_;
    }

private:
    void loop()
    {
        cudaSetDevice(device_);
        for (;;) {
            std::function<int()> job;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [this] { return hasJob_ || stop_; });
                if (stop_) return;
                job = std::move(job_);
                hasJob_ = false;
            }
            const int rc = job();
            {
                std::lock_guard<std::mutex> g(m_);
                rc_ = rc;
                done_ = true;
            }
            cv_.notify_all();
        }
    }
    int device_;
    std::mutex m_;
    std::condition_variable cv_;
    std::function<int()> job_;
    bool hasJob_ = false, done_ = true, stop_ = false;
    int rc_ = 0;
    std::thread thread_;
};

std::mutex g_schedMutex;  // one sharded call at a time per process
std::vector<DeviceWorker*> g_workers;  // leaked at exit on purpose (see HostPipe)

int device_count()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

DeviceWorker* worker_for(int device)
{
    if ((int)g_workers.size() <= device) g_workers.resize((size_t)device + 1, nullptr);
    if (!g_workers[(size_t)device]) g_workers[(size_t)device] = new DeviceWorker(device);
    return g_workers[(size_t)device];
}

}  // namespace

// ======================================================================== extern "C"
extern "C" {

int goofy_b200_abi_version(void) { return GOOFY_B200_ABI_VERSION; }

int goofy_b200_device_count(void) { return device_count(); }

uint64_t goofy_b200_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

int goofy_b200_set_load_path(int path)
{
    if (path < GOOFY_B200_LOAD_AUTO || path > GOOFY_B200_LOAD_ASYNC) return GOOFY_B200_E_ARGS;
    return g_loadPath.exchange(path, std::memory_order_relaxed);
}

int goofy_b200_get_load_path(void) { return g_loadPath.load(std::memory_order_relaxed); }

const char* goofy_b200_error_string(int code)
{
    switch (code) {
        case GOOFY_B200_OK: return "ok";
        case GOOFY_B200_E_WIDTH: return "width is not a multiple of 16";
        case GOOFY_B200_E_HEIGHT: return "height is not a multiple of 4";
        case GOOFY_B200_E_NULL: return "null pointer";
        case GOOFY_B200_E_ALIGN: return "input/stride must be 16-byte aligned, output 8-byte aligned";
        case GOOFY_B200_E_STRIDE: return "stride smaller than width*4";
        case GOOFY_B200_E_CODEC: return "unknown codec";
        case GOOFY_B200_E_DEVICE: return "no usable CUDA device";
        case GOOFY_B200_E_ARGS: return "invalid argument";
        default: break;
    }
    if (code <= GOOFY_B200_E_CUDA_BASE) return cudaGetErrorString((cudaError_t)(GOOFY_B200_E_CUDA_BASE - code));
    return "unknown error";
}

int goofy_b200_compress_dxt1(unsigned char* result, const unsigned char* input, unsigned int width, unsigned int height,
                             unsigned int stride)
{
    return encode_host(GOOFY_B200_DXT1, result, input, width, height, stride);
}

int goofy_b200_compress_etc1(unsigned char* result, const unsigned char* input, unsigned int width, unsigned int height,
                             unsigned int stride)
{
    return encode_host(GOOFY_B200_ETC1, result, input, width, height, stride);
}

int goofy_b200_encode_host(int codec, void* result, const void* input, uint32_t width, uint32_t height, uint32_t stride)
{
    return encode_host(codec, result, input, width, height, stride);
}

int goofy_b200_encode_device(int codec, void* d_result, const void* d_input, uint32_t width, uint32_t height,
                             uint32_t stride, void* stream)
{
    return encode_any(codec, d_result, d_input, width, height, stride, 0, 0, 1, (cudaStream_t)stream);
}

int goofy_b200_compress_dxt1_floatref(unsigned char* result, const unsigned char* input, unsigned int width,
                                      unsigned int height, unsigned int stride)
{
    return encode_host(GOOFY_B200_DXT1_FLOATREF, result, input, width, height, stride);
}

int goofy_b200_compress_etc1_floatref(unsigned char* result, const unsigned char* input, unsigned int width,
                                      unsigned int height, unsigned int stride)
{
    return encode_host(GOOFY_B200_ETC1_FLOATREF, result, input, width, height, stride);
}

int goofy_b200_encode_batch_uniform_device(int codec, void* d_result, const void* d_input, uint32_t width, uint32_t height,
                                           uint32_t stride, uint64_t input_image_pitch, uint64_t result_image_pitch,
                                           uint32_t n_images, void* stream)
{
    return encode_any(codec, d_result, d_input, width, height, stride, input_image_pitch, result_image_pitch, n_images,
                      (cudaStream_t)stream);
}

int goofy_b200_encode_dual_device(void* d_result_dxt1, void* d_result_etc1, const void* d_input, uint32_t width,
                                  uint32_t height, uint32_t stride, uint64_t input_image_pitch, uint64_t result_image_pitch,
                                  uint32_t n_images, void* stream)
{
    return encode_uniform(gb::kDual, d_result_dxt1, d_result_etc1, d_input, width, height, stride, input_image_pitch,
                          result_image_pitch, n_images, (cudaStream_t)stream);
}

int goofy_b200_encode_relaxed_device(int codec, void* d_result, const void* d_input, uint32_t width, uint32_t height,
                                     uint32_t stride, void* stream)
{
    if (!is_codec(codec)) return GOOFY_B200_E_CODEC;
    if (width == 0u || height == 0u) return GOOFY_B200_OK;
    if ((uint64_t)stride < (uint64_t)width * 4u) return GOOFY_B200_E_STRIDE;
    if (!d_result || !d_input) return GOOFY_B200_E_NULL;
    if (((uintptr_t)d_input & 3u) != 0u || (stride & 3u) != 0u || ((uintptr_t)d_result & 7u) != 0u) return GOOFY_B200_E_ALIGN;
    const uint32_t bw = (width + 3u) / 4u, bh = (height + 3u) / 4u;
    if (bh > 65535u) return GOOFY_B200_E_ARGS;
    int rc = ensure_device_ready();
    if (rc != GOOFY_B200_OK) return rc;
    const dim3 grid((bw + 255u) / 256u, bh, 1);
    cudaStream_t s = (cudaStream_t)stream;
    const uint8_t* src = (const uint8_t*)d_input;
    uint8_t* dst = (uint8_t*)d_result;
    switch (codec) {
        case GOOFY_B200_DXT1: gb::encode_relaxed_kernel<gb::kDxt1, 0><<<grid, 256, 0, s>>>(src, dst, width, height, stride); break;
        case GOOFY_B200_ETC1: gb::encode_relaxed_kernel<gb::kEtc1, 0><<<grid, 256, 0, s>>>(src, dst, width, height, stride); break;
        case GOOFY_B200_DXT1_FLOATREF: gb::encode_relaxed_kernel<gb::kDxt1, 1><<<grid, 256, 0, s>>>(src, dst, width, height, stride); break;
        default: gb::encode_relaxed_kernel<gb::kEtc1, 1><<<grid, 256, 0, s>>>(src, dst, width, height, stride); break;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(cudaGetLastError());
}

int goofy_b200_decode_device(int codec, void* d_rgba, const void* d_blocks, uint32_t width, uint32_t height, uint32_t stride,
                             void* stream)
{
    if (codec != GOOFY_B200_DXT1 && codec != GOOFY_B200_ETC1) return GOOFY_B200_E_CODEC;
    if (width % 4u != 0u) return GOOFY_B200_E_WIDTH;
    if (height % 4u != 0u) return GOOFY_B200_E_HEIGHT;
    if (width == 0u || height == 0u) return GOOFY_B200_OK;
    if ((uint64_t)stride < (uint64_t)width * 4u) return GOOFY_B200_E_STRIDE;
    if (!d_rgba || !d_blocks) return GOOFY_B200_E_NULL;
    if (((uintptr_t)d_rgba & 15u) != 0u || (stride & 15u) != 0u || ((uintptr_t)d_blocks & 7u) != 0u) return GOOFY_B200_E_ALIGN;
    if (height / 4u > 65535u) return GOOFY_B200_E_ARGS;
    int rc = ensure_device_ready();
    if (rc != GOOFY_B200_OK) return rc;
    gb::DecodeParams P;
    P.blocks = (const uint8_t*)d_blocks;
    P.rgba = (uint8_t*)d_rgba;
    P.source = nullptr;
    P.sse = nullptr;
    P.bw = width / 4u;
    P.bh = height / 4u;
    P.stride = stride;
    const dim3 grid((P.bw + 255u) / 256u, P.bh, 1);
    if (codec == GOOFY_B200_DXT1) gb::decode_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    else gb::decode_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(cudaGetLastError());
}

int goofy_b200_block_sse_device(int codec, const void* d_blocks, const void* d_rgba, uint32_t width, uint32_t height,
                                uint32_t stride, uint64_t* d_sse_rgb, void* stream)
{
    if (codec != GOOFY_B200_DXT1 && codec != GOOFY_B200_ETC1) return GOOFY_B200_E_CODEC;
    if (width % 4u != 0u) return GOOFY_B200_E_WIDTH;
    if (height % 4u != 0u) return GOOFY_B200_E_HEIGHT;
    if (width == 0u || height == 0u) return GOOFY_B200_OK;
    if ((uint64_t)stride < (uint64_t)width * 4u) return GOOFY_B200_E_STRIDE;
    if (!d_rgba || !d_blocks || !d_sse_rgb) return GOOFY_B200_E_NULL;
    if (((uintptr_t)d_rgba & 15u) != 0u || (stride & 15u) != 0u || ((uintptr_t)d_blocks & 7u) != 0u ||
        ((uintptr_t)d_sse_rgb & 7u) != 0u)
        return GOOFY_B200_E_ALIGN;
    if (height / 4u > 65535u) return GOOFY_B200_E_ARGS;
    int rc = ensure_device_ready();
    if (rc != GOOFY_B200_OK) return rc;
    gb::DecodeParams P;
    P.blocks = (const uint8_t*)d_blocks;
    P.rgba = nullptr;
    P.source = (const uint8_t*)d_rgba;
    P.sse = (unsigned long long*)d_sse_rgb;
    P.bw = width / 4u;
    P.bh = height / 4u;
    P.stride = stride;
    const dim3 grid((P.bw + 255u) / 256u, P.bh, 1);
    if (codec == GOOFY_B200_DXT1) gb::block_sse_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    else gb::block_sse_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(cudaGetLastError());
}

int goofy_b200_encode_batch_device(int codec, const GoofyB200Image* descs, uint32_t n_images, void* stream)
{
    if (n_images && descs) {
        int dev = -1;
        if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return GOOFY_B200_E_DEVICE; }
        for (uint32_t i = 0; i < n_images; ++i)
            if (descs[i].device >= 0 && descs[i].device != dev) return GOOFY_B200_E_DEVICE;
    }
    return encode_batch_current_device(codec, descs, nullptr, n_images, (cudaStream_t)stream);
}

int goofy_b200_encode_batch_sharded(int codec, const GoofyB200Image* descs, uint32_t n_images)
{
    if (codec != GOOFY_B200_DXT1 && codec != GOOFY_B200_ETC1) return GOOFY_B200_E_CODEC;
    if (n_images == 0u) return GOOFY_B200_OK;
    if (!descs) return GOOFY_B200_E_NULL;
    const int nDev = device_count();
    if (nDev <= 0) return GOOFY_B200_E_DEVICE;
    std::vector<std::vector<uint32_t>> perDevice((size_t)nDev);
    for (uint32_t i = 0; i < n_images; ++i) {
        if (descs[i].device < 0 || descs[i].device >= nDev) return GOOFY_B200_E_DEVICE;
        perDevice[(size_t)descs[i].device].push_back(i);
    }
    std::lock_guard<std::mutex> lock(g_schedMutex);
    std::vector<int> used;
    for (int d = 0; d < nDev; ++d) {
        if (perDevice[(size_t)d].empty()) continue;
        const std::vector<uint32_t>* order = &perDevice[(size_t)d];
        worker_for(d)->submit([codec, descs, order]() {
            int rc = encode_batch_current_device(codec, descs, order->data(), (uint32_t)order->size(), nullptr);
            if (rc != GOOFY_B200_OK) return rc;
            return cuda_rc(cudaStreamSynchronize(nullptr));
        });
        used.push_back(d);
    }
    int rc = GOOFY_B200_OK;
    for (int d : used) {
        const int r = g_workers[(size_t)d]->wait();
        if (r != GOOFY_B200_OK && rc == GOOFY_B200_OK) rc = r;
    }
    return rc;
}

void goofy_b200_strip_partition(uint32_t height, int n_shards, int shard, uint32_t* first_block_row,
                                uint32_t* block_row_count)
{
    const uint32_t rows = height / 4u;
    uint32_t first = 0, count = 0;
    if (n_shards > 0 && shard >= 0 && shard < n_shards) {
        first = (uint32_t)((uint64_t)rows * (uint32_t)shard / (uint32_t)n_shards);
        const uint32_t next = (uint32_t)((uint64_t)rows * ((uint32_t)shard + 1u) / (uint32_t)n_shards);
        count = next - first;
    }
    if (first_block_row) *first_block_row = first;
    if (block_row_count) *block_row_count = count;
}

int goofy_b200_encode_sharded_host(int codec, void* result, const void* input, uint32_t width, uint32_t height,
                                   uint32_t stride, int n_gpus)
{
    if (codec != GOOFY_B200_DXT1 && codec != GOOFY_B200_ETC1) return GOOFY_B200_E_CODEC;
    int rc = check_shape(width, height, stride);
    if (rc != GOOFY_B200_OK) return rc;
    if (width == 0u || height == 0u) return GOOFY_B200_OK;
    if (!result || !input) return GOOFY_B200_E_NULL;
    const int nDev = device_count();
    if (nDev <= 0) return GOOFY_B200_E_DEVICE;
    if (n_gpus <= 0) n_gpus = nDev;
    if (n_gpus > nDev) return GOOFY_B200_E_DEVICE;

    std::lock_guard<std::mutex> lock(g_schedMutex);
    std::vector<int> used;
    for (int g = 0; g < n_gpus; ++g) {
        uint32_t first = 0, count = 0;
        goofy_b200_strip_partition(height, n_gpus, g, &first, &count);
        if (count == 0u) continue;
        uint8_t* out = (uint8_t*)result + (size_t)first * (width / 4u) * 8u;
        const uint8_t* in = (const uint8_t*)input + (size_t)first * 4u * stride;
        worker_for(g)->submit([=]() { return encode_host(codec, out, in, width, count * 4u, stride); });
        used.push_back(g);
    }
    for (int g : used) {
        const int r = g_workers[(size_t)g]->wait();
        if (r != GOOFY_B200_OK && rc == GOOFY_B200_OK) rc = r;
    }
    return rc;
}

}  // extern "C"
