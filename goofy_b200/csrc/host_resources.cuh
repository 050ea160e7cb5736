// host_resources.cuh -- the scratch a host thread needs for the host-pointer path and for ragged batches: streams,
// device strips, pinned staging strips, the descriptor arena.
//
// Lifetime: the sets live in a process-wide pool and a thread only LEASES one (a thread_local handle whose destructor
// hands the set back -- no CUDA call in it, so it is safe however late it runs).  A caller that replaces the
// multi-threaded CPU reference with short-lived threads (std::async, a thread per texture) therefore recycles a
// handful of sets -- as many as there were threads inside the library at the same time -- instead of leaking about
// 100 MB of device and pinned memory per thread that ever called.  The pool itself is never destroyed (static
// destruction can run after the CUDA runtime has shut down; the driver reclaims everything at process exit).
#pragma once
#include "host_launch.cuh"

namespace {

// ---------------------------------------------------------------- host-pointer pipeline
// The image is cut into strips of whole block rows; strip i runs H2D -> kernel -> D2H on
// stream i % kSlots so the copies of neighbouring strips overlap each other and the kernels.
// Device scratch only ever grows.
constexpr int kSlots = 3;
constexpr size_t kStripBytes = 16u << 20;
constexpr int kFlights = 24;     // hybrid pinned path: uploads that may be queued on the link at once
constexpr int kPackSlots = 4;    // ... and pinned strips that hold alpha-stripped pixels on their way to the device

struct HostPipe {
    int device = -1;
    cudaStream_t stream[kSlots] = {};
    void* dIn[kSlots] = {};
    void* dOut[kSlots] = {};
    size_t capIn = 0, capOut = 0;
    bool ready = false;
    cudaEvent_t uploaded[kFlights] = {};   // created on first use by the hybrid pinned path

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
    int ensure_events()
    {
        for (int i = 0; i < kFlights; ++i)
            if (!uploaded[i]) GB_CUDA(cudaEventCreateWithFlags(&uploaded[i], cudaEventDisableTiming));
        return GOOFY_B200_OK;
    }
    void release()   // the set moves to another device: its streams and strips belong to the old one
    {
        for (int i = 0; i < kSlots; ++i) {
            if (dIn[i]) cudaFree(dIn[i]);
            if (dOut[i]) cudaFree(dOut[i]);
            if (stream[i]) cudaStreamDestroy(stream[i]);
            dIn[i] = dOut[i] = nullptr;
            stream[i] = nullptr;
        }
        for (int i = 0; i < kFlights; ++i) {
            if (uploaded[i]) cudaEventDestroy(uploaded[i]);
            uploaded[i] = nullptr;
        }
        cudaGetLastError();
        capIn = capOut = 0;
        ready = false;
    }
};

// Pinned staging strips, allocated only when a pageable buffer is first seen.
struct HostStage {
    void* in[kSlots] = {};
    void* out[kSlots] = {};
    void* pack[kPackSlots] = {};   // hybrid pinned path: alpha-stripped strips
    size_t capIn = 0, capOut = 0, capPack = 0;
    int ensure_pack(size_t need)
    {
        if (need > capPack) {
            capPack = 0;
            for (int i = 0; i < kPackSlots; ++i) {
                if (pack[i]) cudaFreeHost(pack[i]);
                pack[i] = nullptr;
                GB_CUDA(cudaHostAlloc(&pack[i], need, cudaHostAllocDefault));
            }
            capPack = need;
        }
        return GOOFY_B200_OK;
    }
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

// Descriptor tables of large ragged batches: a small device arena recycled in stream order.
struct BatchArena {
    void* dev = nullptr;
    void* host = nullptr;  // pinned
    size_t cap = 0;
    int device = -1;
    cudaEvent_t done = nullptr;   // belongs to `device`, like `dev`

    // Make room for `bytes` on device `dev`; waits until the previous batch has consumed the table.
    int prepare(int devNow, size_t bytes)
    {
        if (done) GB_CUDA(cudaEventSynchronize(done));
        if (device == devNow && bytes <= cap) return GOOFY_B200_OK;
        // other device or too small: everything device-bound goes, the event included (an event recorded on a
        // stream of another device is cudaErrorInvalidResourceHandle -- after the kernel has been launched)
        if (dev) cudaFree(dev);
        if (host) cudaFreeHost(host);
        if (done) cudaEventDestroy(done);
        cudaGetLastError();
        dev = host = nullptr;
        done = nullptr;
        cap = 0;
        device = -1;
        const size_t want = bytes < (1u << 16) ? (1u << 16) : bytes * 2u;
        GB_CUDA(cudaMalloc(&dev, want));
        GB_CUDA(cudaHostAlloc(&host, want, cudaHostAllocDefault));
        GB_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        cap = want;
        device = devNow;
        return GOOFY_B200_OK;
    }
};

struct ThreadResources {
    HostPipe pipe;
    HostStage stage;
    BatchArena arena;
};

class ResourcePool {
public:
    static ResourcePool& get()
    {
        static ResourcePool* pool = new ResourcePool();   // leaked on purpose, see the top of this file
        return *pool;
    }
    ThreadResources* take(int preferDevice)
    {
        std::lock_guard<std::mutex> g(m_);
        for (size_t i = 0; i < idle_.size(); ++i)
            if (idle_[i]->pipe.device == preferDevice || i + 1 == idle_.size()) {
                ThreadResources* r = idle_[i];
                idle_.erase(idle_.begin() + (long)i);
                return r;
            }
        created_.fetch_add(1, std::memory_order_relaxed);
        return new ThreadResources();
    }
    void give_back(ThreadResources* r)
    {
        std::lock_guard<std::mutex> g(m_);
        idle_.push_back(r);
    }
    uint64_t created() const { return created_.load(std::memory_order_relaxed); }

private:
    std::mutex m_;
    std::vector<ThreadResources*> idle_;
    std::atomic<uint64_t> created_{0};
};

struct ResourceLease {
    ThreadResources* r = nullptr;
    ~ResourceLease()
    {
        if (r) ResourcePool::get().give_back(r);   // no CUDA call: safe even during process teardown
    }
};
thread_local ResourceLease t_lease;

ThreadResources& thread_resources(int device)
{
    if (!t_lease.r) t_lease.r = ResourcePool::get().take(device);
    return *t_lease.r;
}

}  // namespace
