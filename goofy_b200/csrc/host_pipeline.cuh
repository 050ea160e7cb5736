// host_pipeline.cuh -- the drop-in host-pointer path: strips of block rows pipelined H2D -> kernel -> D2H over
// three streams; pageable buffers staged through pinned strips by a small pool of copy threads.
#pragma once
#include "host_resources.cuh"
#include "copy_pool.h"
#include "hybrid_choice.h"
#include "host_neighbours.cuh"

namespace {

// Is this host pointer ordinary (pageable) memory?  cudaPointerGetAttributes costs about a microsecond, twice per call,
// which a 50 us call on a test-image-sized texture notices; harnesses call with the same buffers over and over
// (the reference's: 128 times per image, Src/main.cpp:653), so the last few answers are remembered per thread.
// A stale answer (the range was freed and came back as the other kind) is harmless: pinned memory can be staged
// like pageable memory, and cudaMemcpyAsync accepts pageable memory (the driver then stages it itself).
bool is_pageable(const void* p)
{
    struct Entry { const void* p; bool pageable; };
    thread_local Entry cache[4] = {};
    thread_local unsigned next = 0;
    for (const Entry& e : cache)
        if (e.p == p && p) return e.pageable;
    cudaPointerAttributes a;
    bool pageable = true;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) cudaGetLastError();
    else pageable = a.type == cudaMemoryTypeUnregistered;
    cache[next++ & 3u] = Entry{p, pageable};
    return pageable;
}

// Alpha-stripped staging (rgb_pack.h): 0 = never, 1 = AUTO, 2 = every strip of a pinned image too (experiments),
// 3 = pageable input only (what a launcher that runs one process per GPU on a shared host should set).
// AUTO: pageable input is staged as packed RGB (the staging copy has to be made anyway; it then writes, and the link
// then carries, 3 bytes per pixel instead of 4); large pinned input goes through the hybrid scheduler (run_hybrid).
std::atomic<int> g_hostRgb{-1};
int host_rgb_mode()
{
    int m = g_hostRgb.load(std::memory_order_relaxed);
    if (m < 0) {
        m = env_int("GOOFY_B200_HOST_RGB", 0, 3, 1);
        g_hostRgb.store(m, std::memory_order_relaxed);
    }
    return m;
}
// bytes that crossed the link host -> device through the host path (copy engine or zero-copy kernel reads), and
// strips of pinned images sent as they were / alpha-stripped first: what a benchmark reports instead of assuming
std::atomic<uint64_t> g_hostUploaded{0}, g_rawStrips{0}, g_packedStrips{0}, g_packingCalls{0}, g_plainCalls{0};

constexpr size_t kStreamingPackBytes = 32u << 20;   // images beyond this are packed with non-temporal stores (copy_pool.h: pack2d)
constexpr size_t kStagePieceBytes = 768u << 10;   // zero-copy path: smallest piece of a strip staged and encoded on its own
constexpr size_t kCopyPieceBytes = 2u << 20;      // copying path: smallest piece of a strip staged and sent on its own

// GOOFY_B200_TRACE_HOST=n: the n-th host-pointer call of a thread prints where its time went (microseconds since the
// call started, to stderr).  Diagnostic for the small-image path (profiles/r02_hostlat_*.txt); costs one branch otherwise.
struct HostTrace {
    static int wanted() { static const int n = env_int("GOOFY_B200_TRACE_HOST", 1, 1 << 30, 0); return n; }
    std::chrono::steady_clock::time_point t0;
    struct Mark { const char* what; double us; };
    Mark marks[64];
    int n = 0;
    bool on = false;
    void begin(bool enable) { on = enable; n = 0; if (on) t0 = std::chrono::steady_clock::now(); }
    void mark(const char* what)
    {
        if (on && n < 64) marks[n++] = Mark{what, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count()};
    }
    void end()
    {
        if (!on) return;
        for (int i = 0; i < n; ++i) std::fprintf(stderr, "[goofy_b200 host trace] %8.1f us  %s\n", marks[i].us, marks[i].what);
        on = false;
    }
};
thread_local HostTrace t_trace;
thread_local int t_hostCalls = 0;

inline void cpu_pause()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}

// One host image of a host-pointer call, validated and cut into strips.
struct HostJob {
    const uint8_t* input = nullptr;
    uint8_t* result = nullptr;
    uint8_t* result2 = nullptr;   // dual-output jobs: ETC1s blocks (result holds the DXT1 blocks); null otherwise
    bool stageOut2 = false;
    uint32_t width = 0, stride = 0, blockRows = 0, stripRows = 0;
    size_t rowBytes = 0, outRowBytes = 0;
    bool stageIn = false, stageOut = false;
    bool packIn = false;      // pageable input staged as packed RGB (3 bytes per pixel) and encoded by the rgb24 kernels
    bool rgbSource = false;   // the caller's pixels ARE packed RGB (goofy_b200_encode_rgb24_host): nothing to strip
    bool rgbKernels() const { return packIn || rgbSource; }
    size_t stagedRowBytes = 0;   // bytes per pixel row as staged / uploaded: rowBytes, or width * 3 when packIn
};

// Strips of whole block rows: about eight per image so that staging, H2D, kernel and D2H of neighbouring strips
// overlap even for images of a few megabytes, between 2 MiB (below that the per-strip fixed costs dominate: a
// 768x512 image took 74 us in three strips against 52 us in one) and kStripBytes (the scratch a host thread keeps
// per slot; 16 MiB and 8 MiB strips measure the same on 8192^2, 4 MiB and 2 MiB are 30-40 % slower for pageable buffers).
uint32_t strip_rows_for(size_t rowBytes, uint32_t height)
{
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
    const uint32_t blockRows = height / 4u;
    return stripRows > blockRows ? blockRows : stripRows;
}

// The argument checks of one host image, in the reference's order (goofy_tc.h:1500-1508) and then the new ones.
// Returns GOOFY_B200_OK with job.blockRows == 0 for an empty image.
int make_host_job(int codec, void* result, const void* input, uint32_t width, uint32_t height, uint32_t stride, HostJob& job,
                  bool rgbSource = false)
{
    int rc = GOOFY_B200_OK;
    if (rgbSource) {   // packed RGB rows: the codec's own shape rules, 4-byte-aligned rows of at least width * 3 bytes
        if (width % (is_floatref(codec) ? 4u : 16u) != 0u) return GOOFY_B200_E_WIDTH;
        if (height % 4u != 0u) return GOOFY_B200_E_HEIGHT;
        if (width != 0u && height != 0u && (uint64_t)stride < (uint64_t)width * 3u) return GOOFY_B200_E_STRIDE;
        if (width != 0u && height != 0u && (stride & 3u) != 0u) return GOOFY_B200_E_ALIGN;
    } else {
        rc = is_floatref(codec) ? check_shape_floatref(width, height, stride) : check_shape(width, height, stride);
    }
    if (rc != GOOFY_B200_OK) return rc;
    job = HostJob();
    if (width == 0u || height == 0u) return GOOFY_B200_OK;
    if (!result || !input) return GOOFY_B200_E_NULL;
    if (((uintptr_t)input & (rgbSource ? 3u : 15u)) != 0u) return GOOFY_B200_E_ALIGN;  // RGBA: the reference's aligned-load contract
    job.input = (const uint8_t*)input;
    job.result = (uint8_t*)result;
    job.width = width;
    job.stride = stride;
    job.rgbSource = rgbSource;
    job.rowBytes = (size_t)width * (rgbSource ? 3u : 4u);
    job.outRowBytes = (size_t)(width / 4u) * 8u;
    job.blockRows = height / 4u;
    job.stripRows = strip_rows_for(job.rowBytes, height);
    // Pinned buffers are DMA'd in place; pageable ones go through the pinned staging strips.
    job.stageIn = is_pageable(input);
    job.stageOut = is_pageable(result);
    job.packIn = !rgbSource && job.stageIn && host_rgb_mode() != 0;
    job.stagedRowBytes = job.packIn ? (size_t)width * 3u : job.rowBytes;
    return GOOFY_B200_OK;
}

// Large PINNED input: the hybrid scheduler.  The call is bound by the link (4 of the 4.5 bytes per pixel that cross it
// are input, a quarter of them alpha bytes nobody reads), and the host's cores idle while it runs.  So strips are
// claimed from both ends of the image: from the FRONT they are DMA'd as they are; from the BACK the calling thread and
// the copy pool first drop the alpha byte (CopyPool::pack2d, into a small ring of pinned strips) and 3 bytes per pixel
// cross instead of 4, encoded by the rgb24 kernels.  The two kinds meet wherever the host's packing rate puts them: the
// caller keeps enough raw uploads queued to cover the time one pack takes (measured, a running mean) and packs only
// while that much is queued, so a host with no cores to spare degrades to the plain DMA pipeline and a host with
// plenty sends nearly everything packed.  With packing rate P and link rate L (input bytes per second) the share of
// packed strips settles at x = 1 / (L / P + 1/4) (capped at 1) and the call takes x / P: 0.8 of the plain time at
// P = L, 0.75 from P = 4/3 L on.
//
// That holds while the LINK is what bounds the call.  With one process per GPU on a shared host the host's memory
// system can be the bound instead (an 8-GPU box delivers 27-35 GB/s per GPU with four ranks uploading, well below the
// link rate), and packing, which moves 2.5 bytes through host memory per input byte where plain DMA moves one, then
// makes everybody slower (4 ranks: -12...-18 %, sessions S and T).  A process cannot find that out by comparing its own
// calls with and without packing: the cost of its packing falls on the OTHER ranks, so each of them measures packing
// as the better reply to what the others do, and all of them settle on the worse state (tried: session T).  What a
// process can see is the rate of its own PLAIN uploads: a B200 link (PCIe Gen5 x16) carries a plain call at 53-54
// GB/s of input when nothing else holds it back, so AUTO packs only while its plain calls reach kHybridMinPlainGBs (48 GB/s)
// (GOOFY_B200_HYBRID_MIN_LINK_GBS) and -- belt and braces -- while calls that pack measure faster than calls that do
// not.  Every thread's first two calls are plain (the first one cold and not recorded); later every sixteenth call
// runs the way that is NOT preferred to keep both means current.  The bytes produced are the same either way.
thread_local HybridChoice t_hybridChoice{(double)env_int("GOOFY_B200_HYBRID_MIN_LINK_GBS", 0, 1000, kHybridMinPlainGBs) * 1e9};

int run_hybrid(int codec, const HostJob& J, ThreadResources& R, int dev)
{
    const int mode = host_rgb_mode();
    // strips: 1/32 of the image, between 4 and 8 MiB (per-strip host work -- five API calls -- against the granularity
    // of the split and of the tail; sessions O-R, profiles/r02_rgb24_sessions.md); GOOFY_B200_HYBRID_STRIP_KB overrides
    static const size_t stripEnv = (size_t)env_int("GOOFY_B200_HYBRID_STRIP_KB", 256, 16384, 0) << 10;
    static const double linkGBs = (double)env_int("GOOFY_B200_HYBRID_LINK_GBS", 1, 1000, 55);   // only sizes the queue
    size_t stripTarget = stripEnv ? stripEnv : (size_t)J.blockRows * 4u * J.rowBytes / 32u;
    if (!stripEnv) stripTarget = stripTarget < (4u << 20) ? (4u << 20) : stripTarget > (8u << 20) ? (8u << 20) : stripTarget;
    uint32_t stripRows = (uint32_t)(stripTarget / (J.rowBytes * 4u));
    if (stripRows == 0u) stripRows = 1u;
    const uint32_t nStrips = (J.blockRows + stripRows - 1u) / stripRows;
    const size_t stripIn = (size_t)stripRows * 4u * J.rowBytes, packedRow = (size_t)J.width * 3u;
    const size_t stripPacked = (size_t)stripRows * 4u * packedRow;
    const size_t half = (size_t)stripRows * J.outRowBytes;
    int rc = R.pipe.prepare(dev, stripIn, half * (J.result2 ? 2u : 1u));
    if (rc != GOOFY_B200_OK) return rc;
    rc = R.pipe.ensure_events();
    if (rc != GOOFY_B200_OK) return rc;
    rc = R.stage.ensure_pack(stripPacked);
    if (rc != GOOFY_B200_OK) return rc;
    const auto callStart = std::chrono::steady_clock::now();   // (after the one-time allocations of a thread's first call)

    struct Flight { bool busy = false; size_t bytes = 0; int packSlot = -1; } flight[kFlights];
    bool packBusy[kPackSlots] = {};
    size_t queuedBytes = 0;
    int nFlights = 0;
    auto poll = [&]() -> int {   // retire uploads that have finished
        for (int i = 0; i < kFlights; ++i) {
            if (!flight[i].busy) continue;
            const cudaError_t e = cudaEventQuery(R.pipe.uploaded[i]);
            if (e == cudaErrorNotReady) {
                cudaGetLastError();   // "not ready" is recorded as the thread's last error: do not leave it for the caller's next check
                continue;
            }
            if (e != cudaSuccess) return cuda_rc(e);
            flight[i].busy = false;
            queuedBytes -= flight[i].bytes;
            --nFlights;
            if (flight[i].packSlot >= 0) packBusy[flight[i].packSlot] = false;
        }
        return GOOFY_B200_OK;
    };
    auto fail = [&](int code) -> int {
        for (int i = 0; i < kSlots; ++i) cudaStreamSynchronize(R.pipe.stream[i]);
        cudaGetLastError();
        return code;
    };
    uint32_t issued = 0;
    // upload (raw, or from pack slot `packSlot`) -> kernel -> download of strip k, on the next stream round-robin
    auto issue = [&](uint32_t k, int packSlot) -> int {
        const int slot = (int)(issued++ % (uint32_t)kSlots);
        cudaStream_t s = R.pipe.stream[slot];
        const uint32_t r0 = k * stripRows, rows = J.blockRows - r0 < stripRows ? J.blockRows - r0 : stripRows;
        int f = 0;
        while (flight[f].busy) ++f;   // the caller made sure one is free
        const size_t bytes = (size_t)rows * 4u * (packSlot >= 0 ? packedRow : J.rowBytes);
        if (packSlot >= 0)
            GB_CUDA(cudaMemcpyAsync(R.pipe.dIn[slot], R.stage.pack[packSlot], bytes, cudaMemcpyHostToDevice, s));
        else
            GB_CUDA(cudaMemcpy2DAsync(R.pipe.dIn[slot], J.rowBytes, J.input + (size_t)r0 * 4u * J.stride, J.stride, J.rowBytes, (size_t)rows * 4u,
                                      cudaMemcpyHostToDevice, s));
        GB_CUDA(cudaEventRecord(R.pipe.uploaded[f], s));
        flight[f].busy = true;
        flight[f].bytes = bytes;
        flight[f].packSlot = packSlot;
        queuedBytes += bytes;
        ++nFlights;
        g_hostUploaded.fetch_add(bytes, std::memory_order_relaxed);
        (packSlot >= 0 ? g_packedStrips : g_rawStrips).fetch_add(1, std::memory_order_relaxed);
        uint8_t* dOut = (uint8_t*)R.pipe.dOut[slot];
        int r;
        if (packSlot >= 0)
            r = encode_rgb24(J.result2 ? GOOFY_B200_BOTH : codec, dOut, dOut + half, R.pipe.dIn[slot], J.width, rows * 4u, (uint32_t)packedRow, 0, 0, 1, s);
        else
            r = J.result2 ? encode_uniform(gb::kDual, dOut, dOut + half, R.pipe.dIn[slot], J.width, rows * 4u, (uint32_t)J.rowBytes, 0, 0, 1, s)
                          : encode_any(codec, dOut, R.pipe.dIn[slot], J.width, rows * 4u, (uint32_t)J.rowBytes, 0, 0, 1, s);
        if (r != GOOFY_B200_OK) return r;
        GB_CUDA(cudaMemcpyAsync(J.result + (size_t)r0 * J.outRowBytes, dOut, (size_t)rows * J.outRowBytes, cudaMemcpyDeviceToHost, s));
        if (J.result2)
            GB_CUDA(cudaMemcpyAsync(J.result2 + (size_t)r0 * J.outRowBytes, dOut + half, (size_t)rows * J.outRowBytes, cudaMemcpyDeviceToHost, s));
        return GOOFY_B200_OK;
    };

    const bool packAll = mode == 2;
    // seconds one pack of a full strip takes: a running mean, seeded per calling thread by its previous call
    thread_local double packSeconds = 0.0;
    if (packSeconds == 0.0) packSeconds = (double)stripIn / 30e9;
    // (A variant in which the calling thread only fed the link while the pool's workers packed asynchronously measured
    // the same, 4420 vs 4422 us on 8192^2, as did 9 or 12 instead of 8 host threads: the packing rate of these hosts is set
    // by their memory system, about 45 GB/s of input, not by the thread count.  Session R.)
    uint32_t front = 0, back = nStrips;
    while (front < back) {
        rc = poll();
        if (rc != GOOFY_B200_OK) return fail(rc);
        // queued upload time that covers one pack, with a margin; never less than two raw strips
        double cover = packSeconds * 1.25 * linkGBs * 1e9;
        if (cover < 2.0 * (double)stripIn) cover = 2.0 * (double)stripIn;
        // the last strips go raw: a strip packed at the very end would add its pack time to the tail of the call, while a
        // raw one only queues behind the uploads already in flight (4096^2: the hybrid call lost to plain DMA without
        // this).  "Last" = what the link drains in the time one more pack would take.
        const bool tail = (double)(back - front) * (double)stripIn <= cover;
        if (!packAll && ((double)queuedBytes < cover || tail) && nFlights < kFlights) {
            rc = issue(front++, -1);
            if (rc != GOOFY_B200_OK) return fail(rc);
            continue;
        }
        int ps = 0;
        while (ps < kPackSlots && packBusy[ps]) ++ps;
        if (ps < kPackSlots && nFlights < kFlights && (packAll || !tail)) {
            const uint32_t k = --back;
            const uint32_t r0 = k * stripRows, rows = J.blockRows - r0 < stripRows ? J.blockRows - r0 : stripRows;
            const auto t0 = std::chrono::steady_clock::now();
            CopyPool::get().pack2d((uint8_t*)R.stage.pack[ps], packedRow, J.input + (size_t)r0 * 4u * J.stride, J.stride, J.width, (size_t)rows * 4u, true);
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * (double)stripRows / (double)rows;
            packSeconds = 0.75 * packSeconds + 0.25 * dt;
            packBusy[ps] = true;
            rc = poll();
            if (rc != GOOFY_B200_OK) return fail(rc);
            rc = issue(k, ps);
            if (rc != GOOFY_B200_OK) return fail(rc);
            continue;
        }
        cpu_pause();
    }
    for (int i = 0; i < kSlots; ++i) {
        const cudaError_t e = cudaStreamSynchronize(R.pipe.stream[i]);
        if (e != cudaSuccess) return fail(cuda_rc(e));
    }
    if (mode != 2)
        t_hybridChoice.record(true, (double)J.blockRows * 4.0 * (double)J.rowBytes /
                                           std::chrono::duration<double>(std::chrono::steady_clock::now() - callStart).count());
    return GOOFY_B200_OK;
}

// The host-pointer pipeline over any number of images: every strip of every image runs stage (if pageable) -> H2D ->
// kernel -> D2H on the stream of its slot, slots taken round-robin, so the copies of one strip (or one small image)
// overlap the kernel and the copies of its neighbours.  Returns when every result is in place.
int run_host_jobs(int codec, const HostJob* jobs, uint32_t nJobs, int dev)
{
    size_t needIn = 0, needOut = 0, needStageIn = 0, needStageOut = 0;
    for (uint32_t j = 0; j < nJobs; ++j) {
        const HostJob& J = jobs[j];
        if (J.blockRows == 0u) continue;
        // dual-output jobs keep their second result in the upper half of the slot's output scratch / staging strip
        const size_t in = (size_t)J.stripRows * 4u * J.stagedRowBytes, out = (size_t)J.stripRows * J.outRowBytes * (J.result2 ? 2u : 1u);
        needIn = in > needIn ? in : needIn;
        needOut = out > needOut ? out : needOut;
        if (J.stageIn && in > needStageIn) needStageIn = in;
        if ((J.stageOut || J.stageOut2) && out > needStageOut) needStageOut = out;
    }
    if (needIn == 0) return GOOFY_B200_OK;
    t_trace.begin(HostTrace::wanted() != 0 && ++t_hostCalls == HostTrace::wanted());
    ThreadResources& R = thread_resources(dev);
    int rc = R.pipe.prepare(dev, needIn, needOut);
    t_trace.mark("scratch ready");
    if (rc != GOOFY_B200_OK) return rc;
    if (needStageIn || needStageOut) {
        rc = R.stage.ensure(needStageIn, needStageOut);
        if (rc != GOOFY_B200_OK) return rc;
    }

    struct Pending { const HostJob* job = nullptr; uint32_t r0 = 0, rows = 0; } pending[kSlots];
    auto retire = [&](int slot) -> int {  // wait for the slot's strip and hand its blocks to the caller
        const HostJob* J = pending[slot].job;
        if (!J) return GOOFY_B200_OK;
        GB_CUDA(cudaStreamSynchronize(R.pipe.stream[slot]));
        t_trace.mark("stream synchronised");
        if (J->stageOut)
            CopyPool::get().copy1d(J->result + (size_t)pending[slot].r0 * J->outRowBytes, (const uint8_t*)R.stage.out[slot],
                                   (size_t)pending[slot].rows * J->outRowBytes);
        if (J->stageOut2)
            CopyPool::get().copy1d(J->result2 + (size_t)pending[slot].r0 * J->outRowBytes,
                                   (const uint8_t*)R.stage.out[slot] + (size_t)J->stripRows * J->outRowBytes,
                                   (size_t)pending[slot].rows * J->outRowBytes);
        pending[slot].job = nullptr;
        t_trace.mark("results copied out of the staging strip");
        return GOOFY_B200_OK;
    };
    // The encode launch of one strip (or one piece of it): both codecs when the job has a second result.
    auto launch = [&](const HostJob& J, uint8_t* out, uint8_t* out2, const uint8_t* in, uint32_t pixelRows, uint32_t stride, cudaStream_t s) -> int {
        if (J.rgbKernels()) return encode_rgb24(J.result2 ? GOOFY_B200_BOTH : codec, out, out2, in, J.width, pixelRows, stride, 0, 0, 1, s);
        return J.result2 ? encode_uniform(gb::kDual, out, out2, in, J.width, pixelRows, stride, 0, 0, 1, s)
                         : encode_any(codec, out, in, J.width, pixelRows, stride, 0, 0, 1, s);
    };
    // stage rows [y0, y1) of a strip into pinned memory: a plain copy, or with the alpha byte dropped on the way
    auto stage_rows = [&](const HostJob& J, uint8_t* stage, const uint8_t* src, size_t rows) {
        if (J.packIn) CopyPool::get().pack2d(stage, J.stagedRowBytes, src, J.stride, J.width, rows, (size_t)J.blockRows * 4u * J.rowBytes > kStreamingPackBytes);
        else CopyPool::get().copy2d(stage, J.rowBytes, src, J.stride, J.rowBytes, rows);
        g_hostUploaded.fetch_add(rows * J.stagedRowBytes, std::memory_order_relaxed);
    };
    // Small strips (a test-image-sized texture is ONE strip of 1.5 MiB) skip the copy engine altogether: the kernel reads
    // the pinned input -- the caller's buffer, or the staging strip -- straight over PCIe and writes its blocks straight
    // into the pinned result (or staging strip).  A 1.5 MiB H2D copy costs 38 us of which 8 are fixed, the 192 KiB D2H copy
    // 14 us of which 10 are fixed, and every extra piece of a piecewise H2D pays the fixed part again (measured with
    // GOOFY_B200_TRACE_HOST, profiles/r02_hostlat_768x512.txt); a kernel launch costs 3-4 us and pulls at link speed.
    // Pageable input is staged in up to four pieces of whole block rows, each encoded by its own launch as soon as it is
    // in pinned memory, so the link works on piece k while the copy threads stage piece k + 1.
    auto issue_zero_copy = [&](int slot, const HostJob& J, uint32_t r0, uint32_t rows, bool& taken) -> int {
        taken = false;
        cudaStream_t s = R.pipe.stream[slot];
        const uint8_t* src = J.input + (size_t)r0 * 4u * J.stride;
        const size_t half = (size_t)J.stripRows * J.outRowBytes;
        void* p = nullptr;
        // device views of the three host buffers the kernel touches; any of them not mapped -> the copying path
        const uint8_t* inDev = nullptr;
        if (J.stageIn) {
            if (cudaHostGetDevicePointer(&p, R.stage.in[slot], 0) != cudaSuccess) { cudaGetLastError(); return GOOFY_B200_OK; }
            inDev = (const uint8_t*)p;
        } else {
            if (cudaHostGetDevicePointer(&p, const_cast<uint8_t*>(src), 0) != cudaSuccess) { cudaGetLastError(); return GOOFY_B200_OK; }
            inDev = (const uint8_t*)p;
        }
        uint8_t *outDev = nullptr, *out2Dev = nullptr;
        void* outHost = J.stageOut ? R.stage.out[slot] : (void*)(J.result + (size_t)r0 * J.outRowBytes);
        if (cudaHostGetDevicePointer(&p, outHost, 0) != cudaSuccess) { cudaGetLastError(); return GOOFY_B200_OK; }
        outDev = (uint8_t*)p;
        if (J.result2) {
            void* out2Host = J.stageOut2 ? (void*)((uint8_t*)R.stage.out[slot] + half) : (void*)(J.result2 + (size_t)r0 * J.outRowBytes);
            if (cudaHostGetDevicePointer(&p, out2Host, 0) != cudaSuccess) { cudaGetLastError(); return GOOFY_B200_OK; }
            out2Dev = (uint8_t*)p;
        }
        if ((((uintptr_t)outDev | (uintptr_t)out2Dev) & 7u) != 0u) return GOOFY_B200_OK;   // the kernels store 8-byte blocks
        taken = true;
        if (!J.stageIn) {
            const int r = launch(J, outDev, out2Dev, inDev, rows * 4u, J.stride, s);
            g_hostUploaded.fetch_add((size_t)rows * 4u * J.rowBytes, std::memory_order_relaxed);
            t_trace.mark("zero-copy kernel launched (pinned input)");
            if (r != GOOFY_B200_OK) return r;
        } else {
            size_t pieces = (size_t)rows * 4u * J.rowBytes / kStagePieceBytes;
            pieces = pieces < 1u ? 1u : (pieces > 4u ? 4u : pieces);
            if (pieces > rows) pieces = rows;
            for (size_t k = 0; k < pieces; ++k) {
                const size_t b0 = (size_t)rows * k / pieces, b1 = (size_t)rows * (k + 1u) / pieces;   // block rows of this piece
                uint8_t* stage = (uint8_t*)R.stage.in[slot] + b0 * 4u * J.stagedRowBytes;
                stage_rows(J, stage, src + b0 * 4u * J.stride, (b1 - b0) * 4u);
                t_trace.mark(J.packIn ? "piece staged into pinned memory, alpha dropped" : "piece staged into pinned memory");
                const int r = launch(J, outDev + b0 * J.outRowBytes, out2Dev ? out2Dev + b0 * J.outRowBytes : nullptr,
                                     inDev + b0 * 4u * J.stagedRowBytes, (uint32_t)(b1 - b0) * 4u, (uint32_t)J.stagedRowBytes, s);
                t_trace.mark("zero-copy kernel launched on the piece");
                if (r != GOOFY_B200_OK) return r;
            }
        }
        pending[slot].job = &J;
        pending[slot].r0 = r0;
        pending[slot].rows = rows;
        return GOOFY_B200_OK;
    };
    static const size_t zeroCopyMax = (size_t)env_int("GOOFY_B200_ZEROCOPY_MAX_KB", 0, 1 << 22, 32768) << 10;
    static const bool zeroCopyPageable = env_int("GOOFY_B200_ZEROCOPY_PAGEABLE", 0, 1, 1) != 0;
    // One strip: stage (if pageable) -> H2D -> kernel -> D2H, all on the slot's stream.
    auto issue = [&](int slot, const HostJob& J, uint32_t r0, uint32_t rows) -> int {
        cudaStream_t s = R.pipe.stream[slot];
        const uint8_t* src = J.input + (size_t)r0 * 4u * J.stride;
        // the staging strips of this slot are about to be reused (pinned results: stream order protects the device scratch)
        if (pending[slot].job && (J.stageIn || J.stageOut || J.stageOut2 || pending[slot].job->stageOut || pending[slot].job->stageOut2)) {
            const int r = retire(slot);
            if (r != GOOFY_B200_OK) return r;
        }
        // decided per IMAGE.  Pinned input: up to 32 MiB (beyond that the copy engine's 53 GB/s beats the ~41 GB/s a kernel
        // pulls over the link: 4096^2 1290 vs 1337 us, 8192^2 5000 vs 5284 us).  Pageable input: always -- the staging copy
        // is the bottleneck there and one step fewer behind it wins at every size (768x512 99 -> 62 us, 2048^2 632 -> 428 us,
        // 8192^2 5940 -> 5524 us; profiles/r02_hostlat_zero_copy.txt).
        if (J.stageIn ? zeroCopyPageable : (size_t)J.blockRows * 4u * J.rowBytes <= zeroCopyMax) {
            bool taken = false;
            const int r = issue_zero_copy(slot, J, r0, rows, taken);
            if (r != GOOFY_B200_OK || taken) return r;
        }
        if (J.stageIn) {
            // staged in up to four pieces, each sent as soon as it is in pinned memory: the DMA of one piece runs under
            // the staging copy of the next (pieces of 2 MiB and more: below that the fixed cost of a copy, about 8 us,
            // outweighs the overlap)
            const size_t pixelRows = (size_t)rows * 4u;
            size_t pieces = pixelRows * J.rowBytes / kCopyPieceBytes;
            pieces = pieces < 1u ? 1u : (pieces > 4u ? 4u : pieces);
            for (size_t k = 0; k < pieces; ++k) {
                const size_t y0 = pixelRows * k / pieces, y1 = pixelRows * (k + 1u) / pieces;
                uint8_t* stage = (uint8_t*)R.stage.in[slot] + y0 * J.stagedRowBytes;
                stage_rows(J, stage, src + y0 * J.stride, y1 - y0);
                t_trace.mark("piece staged into pinned memory");
                GB_CUDA(cudaMemcpyAsync((uint8_t*)R.pipe.dIn[slot] + y0 * J.stagedRowBytes, stage, (y1 - y0) * J.stagedRowBytes, cudaMemcpyHostToDevice, s));
                t_trace.mark("piece H2D issued");
            }
        } else {
            GB_CUDA(cudaMemcpy2DAsync(R.pipe.dIn[slot], J.rowBytes, src, J.stride, J.rowBytes, (size_t)rows * 4u, cudaMemcpyHostToDevice, s));
            g_hostUploaded.fetch_add((size_t)rows * 4u * J.rowBytes, std::memory_order_relaxed);
            t_trace.mark("H2D issued (pinned input)");
        }
        const size_t half = (size_t)J.stripRows * J.outRowBytes;   // where a dual-output job keeps its ETC1s blocks
        uint8_t* dOut = (uint8_t*)R.pipe.dOut[slot];
        const int r = launch(J, dOut, dOut + half, (const uint8_t*)R.pipe.dIn[slot], rows * 4u, (uint32_t)J.stagedRowBytes, s);
        if (r != GOOFY_B200_OK) return r;
        t_trace.mark("kernel launched");
        GB_CUDA(cudaMemcpyAsync(J.stageOut ? R.stage.out[slot] : (void*)(J.result + (size_t)r0 * J.outRowBytes), dOut,
                                (size_t)rows * J.outRowBytes, cudaMemcpyDeviceToHost, s));
        if (J.result2)
            GB_CUDA(cudaMemcpyAsync(J.stageOut2 ? (void*)((uint8_t*)R.stage.out[slot] + half) : (void*)(J.result2 + (size_t)r0 * J.outRowBytes),
                                    dOut + half, (size_t)rows * J.outRowBytes, cudaMemcpyDeviceToHost, s));
        t_trace.mark("D2H issued");
        pending[slot].job = &J;
        pending[slot].r0 = r0;
        pending[slot].rows = rows;
        return GOOFY_B200_OK;
    };
    // On failure nothing may still be reading an input or writing a result when the caller gets control back.
    auto fail = [&](int code) -> int {
        for (int i = 0; i < kSlots; ++i) cudaStreamSynchronize(R.pipe.stream[i]);
        cudaGetLastError();
        return code;
    };

    int slot = 0;
    bool plainTimed = false;
    std::chrono::steady_clock::time_point plainStart;
    for (uint32_t j = 0; j < nJobs; ++j) {
        const HostJob& J = jobs[j];
        // large pinned images (the ones that go through the copy engine strip by strip): with alpha-stripped strips from
        // the back (run_hybrid) when AUTO's measurements say so; otherwise the plain strip pipeline below, timed for them
        if (!J.rgbSource && !J.stageIn && !J.stageOut && !J.stageOut2 && J.blockRows != 0u &&
            (host_rgb_mode() == GOOFY_B200_HOST_RGB_AUTO || host_rgb_mode() == GOOFY_B200_HOST_RGB_ALWAYS) &&
            (size_t)J.blockRows * 4u * J.rowBytes > zeroCopyMax) {
            const bool packing = host_rgb_mode() == 2 || t_hybridChoice.next(Neighbours::count());
            (packing ? g_packingCalls : g_plainCalls).fetch_add(1, std::memory_order_relaxed);
            if (packing) {
                for (int i = 0; i < kSlots; ++i, slot = (slot + 1) % kSlots) {
                    rc = retire(slot);
                    if (rc != GOOFY_B200_OK) return fail(rc);
                }
                rc = run_hybrid(codec, J, R, dev);
                if (rc != GOOFY_B200_OK) return rc;
                t_trace.mark("hybrid pinned image done");
                continue;
            }
            if (nJobs == 1u) {
                plainTimed = true;
                plainStart = std::chrono::steady_clock::now();
            }
        }
        for (uint32_t r0 = 0; r0 < J.blockRows; r0 += J.stripRows, slot = (slot + 1) % kSlots) {
            const uint32_t rows = J.blockRows - r0 < J.stripRows ? J.blockRows - r0 : J.stripRows;
            rc = issue(slot, J, r0, rows);
            if (rc != GOOFY_B200_OK) return fail(rc);
        }
    }
    // drain in issue order (the oldest outstanding strip is in the slot the loop would use next)
    for (int i = 0; i < kSlots; ++i, slot = (slot + 1) % kSlots) {
        rc = retire(slot);
        if (rc != GOOFY_B200_OK) return fail(rc);
    }
    if (plainTimed)
        t_hybridChoice.record(false, (double)jobs[0].blockRows * 4.0 * (double)jobs[0].rowBytes /
                                         std::chrono::duration<double>(std::chrono::steady_clock::now() - plainStart).count());
    t_trace.mark("done");
    t_trace.end();
    return GOOFY_B200_OK;
}

int encode_host(int codec, void* result, const void* input, uint32_t width, uint32_t height, uint32_t stride)
{
    if (!is_codec(codec)) return GOOFY_B200_E_CODEC;
    HostJob job;
    int rc = make_host_job(codec, result, input, width, height, stride, job);
    if (rc != GOOFY_B200_OK || job.blockRows == 0u) return rc;
    int dev = -1;
    rc = ensure_device_ready(&dev);
    if (rc != GOOFY_B200_OK) return rc;
    return run_host_jobs(codec, &job, 1, dev);
}

// A host image whose pixels are packed RGB8 to begin with: 3 bytes per pixel cross the link without any host-side work.
// codec may be GOOFY_B200_BOTH (result2 = the ETC1s blocks).
int encode_rgb24_host(int codec, void* result, void* result2, const void* input, uint32_t width, uint32_t height, uint32_t stride)
{
    const bool both = codec == GOOFY_B200_BOTH;
    if (!both && !is_codec(codec)) return GOOFY_B200_E_CODEC;
    HostJob job;
    int rc = make_host_job(both ? GOOFY_B200_DXT1 : codec, result, input, width, height, stride, job, true);
    if (rc != GOOFY_B200_OK || job.blockRows == 0u) return rc;
    if (both) {
        if (!result2) return GOOFY_B200_E_NULL;
        job.result2 = (uint8_t*)result2;
        job.stageOut2 = is_pageable(result2);
    }
    int dev = -1;
    rc = ensure_device_ready(&dev);
    if (rc != GOOFY_B200_OK) return rc;
    return run_host_jobs(both ? GOOFY_B200_DXT1 : codec, &job, 1, dev);
}

// Both codecs from one upload of a host image: 4 B/px over the link instead of 8.
int encode_dual_host(void* resultDxt1, void* resultEtc1, const void* input, uint32_t width, uint32_t height, uint32_t stride)
{
    HostJob job;
    int rc = make_host_job(GOOFY_B200_DXT1, resultDxt1, input, width, height, stride, job);
    if (rc != GOOFY_B200_OK || job.blockRows == 0u) return rc;
    if (!resultEtc1) return GOOFY_B200_E_NULL;
    job.result2 = (uint8_t*)resultEtc1;
    job.stageOut2 = is_pageable(resultEtc1);
    int dev = -1;
    rc = ensure_device_ready(&dev);
    if (rc != GOOFY_B200_OK) return rc;
    return run_host_jobs(GOOFY_B200_DXT1, &job, 1, dev);
}

// n host images through ONE pipeline: every image is validated before anything is started.
int encode_host_batch(int codec, const GoofyB200Image* images, uint32_t n)
{
    const bool both = codec == GOOFY_B200_BOTH;
    if (!both && !is_codec(codec)) return GOOFY_B200_E_CODEC;
    if (n == 0u) return GOOFY_B200_OK;
    if (!images) return GOOFY_B200_E_NULL;
    std::vector<HostJob> jobs(n);
    for (uint32_t i = 0; i < n; ++i) {
        const int rc = make_host_job(both ? GOOFY_B200_DXT1 : codec, images[i].dst, images[i].src, images[i].width, images[i].height,
                                     images[i].stride, jobs[i]);
        if (rc != GOOFY_B200_OK) return rc;
        if (both && jobs[i].blockRows != 0u) {
            if (!images[i].dst2) return GOOFY_B200_E_NULL;
            jobs[i].result2 = (uint8_t*)images[i].dst2;
            jobs[i].stageOut2 = is_pageable(images[i].dst2);
        }
    }
    int dev = -1;
    const int rc = ensure_device_ready(&dev);
    if (rc != GOOFY_B200_OK) return rc;
    return run_host_jobs(both ? GOOFY_B200_DXT1 : codec, jobs.data(), n, dev);
}

}  // namespace
