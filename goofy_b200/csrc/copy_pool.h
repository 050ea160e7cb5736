// copy_pool.h -- host-side helper of the drop-in host path (host_pipeline.cuh): parallel staging copies between the
// caller's pageable buffers and the library's pinned strips.  Plain C++ (no CUDA), so tests/copy_pool_stress.cpp can
// exercise the wake-up handshake on a machine without a GPU (also under ThreadSanitizer).
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "rgb_pack.h"

namespace {

// Pageable (malloc'd) host buffers cannot be DMA'd directly; the CUDA driver then stages them through
// one small internal buffer at ~10 GB/s.  The library stages them itself instead: a few persistent host
// threads copy each strip into pinned memory in parallel while the previous strips are in flight.
// Workers spin for a few tens of microseconds after a job before they sleep on the condition variable: the
// copies of one call (and of back-to-back calls) then start within a microsecond instead of a futex wake-up and a
// scheduler round trip each (measured on a 768x512 image: 97 us best / 244 us median per call with sleeping workers).
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
        job2d(dst, dstPitch, src, srcPitch, rowBytes, rows, false, false);
    }
    // the same with the alpha byte dropped on the way (rgb_pack.h): src rows of `pixels` RGBA8 pixels -> dst rows of
    // pixels * 3 bytes
    // `streaming`: non-temporal stores.  The packed rows are read next by the DMA engine (or by the GPU over the link),
    // never by this core, and a regular store first reads every destination line into the cache; worth it for images
    // that do not fit the last-level cache anyway (8192^2 pageable 7.0-8.6 -> 6.1-6.6 ms, hybrid pinned 5.5 -> 4.5 ms),
    // not for small ones, whose staged rows the link then reads straight from the cache (2048^2 321 vs 340 us; session P)
    void pack2d(uint8_t* dst, size_t dstPitch, const uint8_t* src, size_t srcPitch, size_t pixels, size_t rows, bool streaming)
    {
        job2d(dst, dstPitch, src, srcPitch, pixels * 4u, rows, true, streaming);
    }
    // threads that work on one job, the caller included
    size_t threads() const { return nWorkers_ + 1; }

    void copy1d(uint8_t* dst, const uint8_t* src, size_t bytes)
    {
        const size_t chunk = 1u << 16, full = bytes / chunk;
        if (full) copy2d(dst, chunk, src, chunk, chunk, full);
        if (bytes > full * chunk) std::memcpy(dst + full * chunk, src + full * chunk, bytes - full * chunk);
    }

private:
    void job2d(uint8_t* dst, size_t dstPitch, const uint8_t* src, size_t srcPitch, size_t rowBytes, size_t rows, bool pack, bool streaming)
    {
        if (rows * rowBytes < kPoolMinBytes || nWorkers_ == 0) {
            run(dst, dstPitch, src, srcPitch, rowBytes, 0, rows, pack, streaming);
            return;
        }
        std::lock_guard<std::mutex> serial(jobMutex_);  // one copy job at a time
        const size_t parts = nWorkers_ + 1;
        publish(dst, dstPitch, src, srcPitch, rowBytes, rows, pack, streaming, parts);
        run(dst, dstPitch, src, srcPitch, rowBytes, rows * (parts - 1) / parts, rows, pack, streaming);  // the caller's share: the last slice
        wait_for_workers();
    }
    // hand rows [0, rows * nWorkers_ / parts) to the workers (jobMutex_ held)
    void publish(uint8_t* dst, size_t dstPitch, const uint8_t* src, size_t srcPitch, size_t rowBytes, size_t rows, bool pack, bool streaming,
                 size_t parts)
    {
        dst_ = dst; dstPitch_ = dstPitch; src_ = src; srcPitch_ = srcPitch; rowBytes_ = rowBytes; rows_ = rows; pack_ = pack; streaming_ = streaming;
        parts_ = parts;
        pending_.store((int)nWorkers_, std::memory_order_relaxed);
        // seq_cst on both sides of the generation_ / sleepers_ handshake (store then load here, store then load in the
        // worker): at least one side sees the other's write, so a worker cannot go to sleep on a published job
        generation_.fetch_add(1);   // publishes the job to spinning workers
        if (sleepers_.load() > 0) {
            { std::lock_guard<std::mutex> g(m_); }               // a worker between its predicate check and its wait
            cv_.notify_all();
        }
    }
    void wait_for_workers()
    {
        for (int spin = 0; pending_.load(std::memory_order_acquire) != 0; ++spin) {
            if (spin < 4096) { cpu_relax(); continue; }
            std::unique_lock<std::mutex> g(m_);
            callerWaiting_ = true;
            done_.wait(g, [this] { return pending_.load(std::memory_order_acquire) == 0; });
            callerWaiting_ = false;
        }
    }

    static constexpr size_t kPoolMinBytes = 256u << 10;   // smaller copies are done by the caller alone
    static constexpr int kSpinMicros = 50;  // how long an idle worker spins before it sleeps

    static void cpu_relax()
    {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    CopyPool()
    {
        unsigned n = std::thread::hardware_concurrency();
        n = n > 16u ? 7u : (n > 2u ? n / 2u - 1u : 0u);  // plus the calling thread
        // GOOFY_B200_HOST_THREADS = threads per staging job, the caller included (1..64): a launcher that runs one
        // process per GPU on a shared host divides the cores between them with it
        if (const char* e = std::getenv("GOOFY_B200_HOST_THREADS")) {
            const int v = std::atoi(e);
            if (v >= 1 && v <= 64) n = (unsigned)v - 1u;
        }
        nWorkers_ = n;
        for (unsigned i = 0; i < n; ++i) std::thread([this, i] { loop(i); }).detach();
    }
    static void run(uint8_t* dst, size_t dstPitch, const uint8_t* src, size_t srcPitch, size_t rowBytes, size_t r0, size_t r1, bool pack, bool streaming)
    {
        if (pack) {
            gbpack::pack_rows(dst, dstPitch, src, srcPitch, rowBytes / 4u, r0, r1, streaming);
            return;
        }
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
            // wait for the next job: spin first, then sleep
            const auto spinUntil = std::chrono::steady_clock::now() + std::chrono::microseconds(kSpinMicros);
            for (unsigned spin = 0; generation_.load(std::memory_order_acquire) == seen; ++spin) {
                if ((spin & 63u) != 63u || std::chrono::steady_clock::now() < spinUntil) { cpu_relax(); continue; }
                std::unique_lock<std::mutex> g(m_);
                sleepers_.fetch_add(1);
                cv_.wait(g, [&] { return generation_.load() != seen; });
                sleepers_.fetch_sub(1);
            }
            seen = generation_.load(std::memory_order_acquire);
            const size_t parts = parts_;
            run(dst_, dstPitch_, src_, srcPitch_, rowBytes_, rows_ * index / parts, rows_ * (index + 1) / parts, pack_, streaming_);
            if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                std::lock_guard<std::mutex> g(m_);   // the caller may have gone to sleep on done_
                if (callerWaiting_) done_.notify_one();
            }
        }
    }
    size_t nWorkers_ = 0;
    std::mutex jobMutex_, m_;
    std::condition_variable cv_, done_;
    // the current job: written by the submitter before generation_ is bumped (release), read by workers after (acquire)
    uint8_t* dst_ = nullptr;
    const uint8_t* src_ = nullptr;
    size_t dstPitch_ = 0, srcPitch_ = 0, rowBytes_ = 0, rows_ = 0, parts_ = 1;
    bool pack_ = false, streaming_ = false;
    std::atomic<int> pending_{0};
    std::atomic<int> sleepers_{0};
    std::atomic<uint64_t> generation_{0};
    bool callerWaiting_ = false;   // guarded by m_
};

}  // namespace
