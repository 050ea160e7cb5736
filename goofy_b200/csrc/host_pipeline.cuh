// host_pipeline.cuh -- the drop-in host-pointer path: strips of block rows pipelined H2D -> kernel -> D2H over
// three streams; pageable buffers staged through pinned strips by a small pool of copy threads.
#pragma once
#include "host_launch.cuh"

namespace {

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

    // One strip: stage (if pageable) -> H2D -> kernel -> D2H, all on the slot's stream.
    auto issue = [&](int slot, uint32_t r0, uint32_t rows) -> int {
        cudaStream_t s = t_pipe.stream[slot];
        const uint8_t* src = (const uint8_t*)input + (size_t)r0 * 4u * stride;
        if (stageIn || stageOut) {
            const int r = retire(slot);  // the staging strips of this slot are about to be reused
            if (r != GOOFY_B200_OK) return r;
        }
        if (stageIn) {
            CopyPool::get().copy2d((uint8_t*)t_stage.in[slot], rowBytes, src, stride, rowBytes, (size_t)rows * 4u);
            GB_CUDA(cudaMemcpyAsync(t_pipe.dIn[slot], t_stage.in[slot], (size_t)rows * 4u * rowBytes, cudaMemcpyHostToDevice, s));
        } else {
            // stream order protects the slot's device scratch: its previous strip finished D2H on the same stream
            GB_CUDA(cudaMemcpy2DAsync(t_pipe.dIn[slot], rowBytes, src, stride, rowBytes, (size_t)rows * 4u, cudaMemcpyHostToDevice, s));
        }
        const int r = encode_any(codec, t_pipe.dOut[slot], t_pipe.dIn[slot], width, rows * 4u, (uint32_t)rowBytes, 0, 0, 1, s);
        if (r != GOOFY_B200_OK) return r;
        GB_CUDA(cudaMemcpyAsync(stageOut ? t_stage.out[slot] : (void*)((uint8_t*)result + (size_t)r0 * outRowBytes), t_pipe.dOut[slot],
                                (size_t)rows * outRowBytes, cudaMemcpyDeviceToHost, s));
        pending[slot].r0 = r0;
        pending[slot].rows = rows;
        pending[slot].live = true;
        return GOOFY_B200_OK;
    };
    // On failure nothing may still be reading `input` or writing `result` when the caller gets control back.
    auto fail = [&](int code) -> int {
        for (int i = 0; i < kSlots; ++i) cudaStreamSynchronize(t_pipe.stream[i]);
        cudaGetLastError();
        return code;
    };

    int slot = 0;
    for (uint32_t r0 = 0; r0 < blockRows; r0 += stripRows, slot = (slot + 1) % kSlots) {
        const uint32_t rows = blockRows - r0 < stripRows ? blockRows - r0 : stripRows;
        rc = issue(slot, r0, rows);
        if (rc != GOOFY_B200_OK) return fail(rc);
    }
    // drain in strip order (the oldest outstanding strip is in the slot the loop would use next)
    for (int i = 0; i < kSlots; ++i, slot = (slot + 1) % kSlots) {
        rc = retire(slot);
        if (rc != GOOFY_B200_OK) return fail(rc);
    }
    return GOOFY_B200_OK;
}

}  // namespace
