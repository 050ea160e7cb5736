// host_neighbours.cuh -- is this process alone on its host?  Asked by the host path before it spends host memory
// bandwidth on alpha-stripping pinned images (host_pipeline.cuh: run_hybrid): with one process per GPU on a shared host
// the packing of one rank slows the uploads of all the others, and no rank can see that in its own timings.
//
// NVML (dlopen'ed, optional) lists the compute processes of EVERY GPU of the box, whatever CUDA_VISIBLE_DEVICES hides
// from this process.  A GPU counts as a neighbour's when it runs more compute processes than this process accounts for
// (one on every device where its primary context is active: cuDevicePrimaryCtxGetState).  No NVML, or any call failing:
// the answer is "cannot tell" (-1) and the host path falls back to its rate gate.
#pragma once
#include <dlfcn.h>

namespace {

class Neighbours {
public:
    // GPUs of this box that run somebody else's compute process; -1 = cannot tell.  Refreshed at most every few seconds
    // (a query costs about a millisecond per GPU), by whichever caller finds the answer stale; the others use the old one.
    static int count()
    {
        static Neighbours self;
        return self.get();
    }
    static int last() { return s_last.load(std::memory_order_relaxed); }

private:
    typedef int (*InitFn)();
    typedef int (*CountFn)(unsigned*);
    typedef int (*ByIndexFn)(unsigned, void**);
    typedef int (*ByBusIdFn)(const char*, void**);
    typedef int (*IndexFn)(void*, unsigned*);
    typedef int (*ProcsFn)(void*, unsigned*, void*);
    typedef CUresult (*CtxStateFn)(CUdevice, unsigned*, int*);

    InitFn init_ = nullptr;
    CountFn count_ = nullptr;
    ByIndexFn byIndex_ = nullptr;
    ByBusIdFn byBusId_ = nullptr;
    IndexFn index_ = nullptr;
    ProcsFn procs_ = nullptr;
    CtxStateFn ctxState_ = nullptr;
    bool ok_ = false;
    std::mutex m_;
    std::chrono::steady_clock::time_point stamp_;
    bool have_ = false;
    static std::atomic<int> s_last;

    Neighbours()
    {
        static const bool off = env_int("GOOFY_B200_NEIGHBOURS", 0, 1, 1) == 0;   // 0: do not ask (experiments)
        if (off) return;
        void* lib = dlopen("libnvidia-ml.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!lib) return;
        init_ = (InitFn)dlsym(lib, "nvmlInit_v2");
        count_ = (CountFn)dlsym(lib, "nvmlDeviceGetCount_v2");
        byIndex_ = (ByIndexFn)dlsym(lib, "nvmlDeviceGetHandleByIndex_v2");
        byBusId_ = (ByBusIdFn)dlsym(lib, "nvmlDeviceGetHandleByPciBusId_v2");
        index_ = (IndexFn)dlsym(lib, "nvmlDeviceGetIndex");
        procs_ = (ProcsFn)dlsym(lib, "nvmlDeviceGetComputeRunningProcesses_v3");
        if (!procs_) procs_ = (ProcsFn)dlsym(lib, "nvmlDeviceGetComputeRunningProcesses_v2");
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuDevicePrimaryCtxGetState", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            ctxState_ = (CtxStateFn)p;
        else
            cudaGetLastError();
        ok_ = init_ && count_ && byIndex_ && byBusId_ && index_ && procs_ && ctxState_ && init_() == 0;
    }

    int get()
    {
        if (!ok_) return -1;
        std::unique_lock<std::mutex> g(m_, std::try_to_lock);
        if (!g.owns_lock()) return s_last.load(std::memory_order_relaxed);   // somebody else is asking right now
        const auto now = std::chrono::steady_clock::now();
        if (have_ && now - stamp_ < std::chrono::seconds(3)) return s_last.load(std::memory_order_relaxed);
        const int n = query();
        s_last.store(n, std::memory_order_relaxed);
        stamp_ = now;
        have_ = true;
        return n;
    }

    int query() const
    {
        unsigned nGpus = 0;
        if (count_(&nGpus) != 0 || nGpus == 0u || nGpus > 64u) return -1;
        // which of the box's GPUs (NVML indices) carry an active primary context of THIS process
        bool mine[64] = {};
        int nCuda = 0;
        if (cudaGetDeviceCount(&nCuda) != cudaSuccess) { cudaGetLastError(); return -1; }
        for (int d = 0; d < nCuda; ++d) {
            unsigned flags = 0;
            int active = 0;
            if (ctxState_((CUdevice)d, &flags, &active) != CUDA_SUCCESS || !active) continue;
            char bus[32] = {};
            if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), d) != cudaSuccess) { cudaGetLastError(); return -1; }
            void* h = nullptr;
            unsigned idx = 0;
            if (byBusId_(bus, &h) != 0 || index_(h, &idx) != 0 || idx >= 64u) return -1;
            mine[idx] = true;
        }
        int others = 0;
        for (unsigned i = 0; i < nGpus; ++i) {
            void* h = nullptr;
            if (byIndex_(i, &h) != 0) return -1;
            unsigned procs = 0;
            const int r = procs_(h, &procs, nullptr);   // 0 with no process; "insufficient size" (7) with the count otherwise
            if (r != 0 && r != 7) return -1;
            if (procs > (mine[i] ? 1u : 0u)) ++others;
        }
        return others;
    }
};
std::atomic<int> Neighbours::s_last{-1};

}  // namespace
