// host_common.cuh -- shared host-side state of libgoofy_b200.so: error mapping, per-device one-time set-up,
// the argument checks that mirror goofy::compressDXT1/ETC1 (GoofyTC/goofy_tc.h:1497-1557).
// Included by capi.cu only (one translation unit: the kernels' __device__ tables live in it).
#pragma once
namespace {

std::atomic<uint64_t> g_launches{0};
// what the calling thread launched last (goofy_b200_last_launch_kernel): set by every launcher, so a benchmark can
// report the kernel a call actually ran instead of guessing it from its flags
thread_local const char* t_lastKernel = "";

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? GOOFY_B200_OK : GOOFY_B200_E_CUDA_BASE - (int)e; }

#define GB_CUDA(call)                                   \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return cuda_rc(e__);    \
    } while (0)

std::atomic<int> g_loadPath{GOOFY_B200_LOAD_AUTO};

// The calling thread's current device.  There is no per-device initialisation: the kernels compute their control
// table themselves, so the first call on a device does not synchronise it and the device entry points can be
// captured into a CUDA graph from their very first use.
constexpr int kMaxDevices = 64;

int ensure_device_ready(int* deviceOut = nullptr)
{
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cudaGetLastError(); return GOOFY_B200_E_DEVICE; }
    if (dev < 0 || dev >= kMaxDevices) return GOOFY_B200_E_DEVICE;
    if (deviceOut) *deviceOut = dev;
    return GOOFY_B200_OK;
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

}  // namespace
