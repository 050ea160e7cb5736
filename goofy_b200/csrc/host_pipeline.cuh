// host_pipeline.cuh -- the drop-in host-pointer path: strips of block rows pipelined H2D -> kernel -> D2H over
// three streams; pageable buffers staged through pinned strips by a small pool of copy threads.
#pragma once
#include "host_launch.cuh"
#include "copy_pool.h"

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
    // Strips of whole block rows: about eight per image so that staging, H2D, kernel and D2H of neighbouring strips
    // overlap even for images of a few megabytes, between 2 MiB (below that the per-strip fixed costs dominate: a
    // 768x512 image took 74 us in three strips against 52 us in one) and kStripBytes (the scratch a host thread keeps
    // per slot; 16 MiB and 8 MiB strips measure the same on 8192^2, 4 MiB and 2 MiB are 30-40 % slower for pageable buffers).
    const size_t totalIn = rowBytes * (size_t)height;
    size_t stripTarget = totalIn / 8u;
    if (stripTarget < (2u << 20)) stripTarget = 2u << 20;
    static const size_t stripCap = []() -> size_t {   // GOOFY_B200_STRIP_MB: experiments only (1..16)
        const char* e = getenv("GOOFY_B200_STRIP_MB");
        const int v = e ? atoi(e) : 0;
        return (v >= 1 && v <= 16) ? (size_t)v << 20 : kStripBytes;
    }();
    if (stripTarget > stripCap) stripTarget = stripCap;
    uint32_t stripRows = (uint32_t)(stripTarget / (rowBytes * 4u));
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
