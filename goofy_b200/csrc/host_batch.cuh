// host_batch.cuh -- ragged batches (images of different shapes in one launch) and the multi-GPU shard
// scheduler: one persistent host thread per device, static partition, no collectives.
#pragma once
#include "host_pipeline.cuh"

namespace {

// ---------------------------------------------------------------- ragged batch
template <int MODE, int FLAVOUR, typename TABLE>
int launch_batch(const TABLE& table, uint32_t n, uint32_t totalCtas, cudaStream_t stream)
{
    t_lastKernel = MODE == gb::kDxt1 ? "encode_batch_kernel<dxt1>" : MODE == gb::kEtc1 ? "encode_batch_kernel<etc1s>" : "encode_batch_kernel<dxt1+etc1s>";
    return launch_pdl(gb::encode_batch_kernel<MODE, FLAVOUR, TABLE>, dim3(totalCtas, 1, 1), dim3(gb::kBatchTileX, gb::kBatchTileY, 1), stream, table, n);
}

// codec: DXT1, ETC1, BOTH, or a float-reference flavour
template <typename TABLE>
int launch_batch_codec(int codec, const TABLE& table, uint32_t n, uint32_t totalCtas, cudaStream_t stream)
{
    switch (codec) {
        case GOOFY_B200_DXT1: return launch_batch<gb::kDxt1, 0>(table, n, totalCtas, stream);
        case GOOFY_B200_ETC1: return launch_batch<gb::kEtc1, 0>(table, n, totalCtas, stream);
        case GOOFY_B200_BOTH: return launch_batch<gb::kDual, 0>(table, n, totalCtas, stream);
        case GOOFY_B200_DXT1_FLOATREF: return launch_batch<gb::kDxt1, 1>(table, n, totalCtas, stream);
        case GOOFY_B200_ETC1_FLOATREF: return launch_batch<gb::kEtc1, 1>(table, n, totalCtas, stream);
        default: return GOOFY_B200_E_CODEC;
    }
}

int encode_batch_current_device(int codec, const GoofyB200Image* descs, const uint32_t* order, uint32_t n, cudaStream_t stream)
{
    const bool both = codec == GOOFY_B200_BOTH;
    if (!both && !is_codec(codec)) return GOOFY_B200_E_CODEC;
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
        rc = is_floatref(codec) ? check_shape_floatref(d.width, d.height, d.stride) : check_shape(d.width, d.height, d.stride);
        if (rc != GOOFY_B200_OK) return rc;
        if (d.width == 0u || d.height == 0u) continue;
        rc = check_pointers(d.src, d.dst);
        if (rc != GOOFY_B200_OK) return rc;
        if (both) {
            rc = check_pointers(d.src, d.dst2);
            if (rc != GOOFY_B200_OK) return rc;
        }
        gb::BatchImage im;
        im.src = (const uint8_t*)d.src;
        im.dst = (uint8_t*)d.dst;
        im.dst2 = both ? (uint8_t*)d.dst2 : nullptr;
        im.bw = d.width / 4u;
        im.bh = d.height / 4u;
        im.stride = d.stride;
        im.tilesX = (im.bw + gb::kBatchTileX - 1u) / gb::kBatchTileX;
        start.push_back((uint32_t)total);
        total += (uint64_t)im.tilesX * ((im.bh + gb::kBatchTileY * gb::kBatchPasses - 1u) / (gb::kBatchTileY * gb::kBatchPasses));
        if (total > 0x7FFFFFFFull) return GOOFY_B200_E_ARGS;
        images.push_back(im);
    }
    if (images.empty()) return GOOFY_B200_OK;
    const uint32_t m = (uint32_t)images.size();
    if (m <= (uint32_t)gb::kBatchInline) {
        // small batch: descriptors and prefix sums go in as kernel parameters (no table upload, no arena)
        gb::BatchTableInline T;
        std::memset(&T, 0, sizeof(T));
        std::memcpy(T.images, images.data(), (size_t)m * sizeof(gb::BatchImage));
        std::memcpy(T.ctaStart, start.data(), (size_t)m * sizeof(uint32_t));
        return launch_batch_codec(codec, T, m, (uint32_t)total, stream);
    }
    const size_t bytesImages = (size_t)m * sizeof(gb::BatchImage);
    const size_t bytes = bytesImages + (size_t)m * sizeof(uint32_t);

    BatchArena& A = thread_resources(dev).arena;
    rc = A.prepare(dev, bytes);
    if (rc != GOOFY_B200_OK) return rc;
    std::memcpy(A.host, images.data(), bytesImages);
    std::memcpy((uint8_t*)A.host + bytesImages, start.data(), (size_t)m * sizeof(uint32_t));
    GB_CUDA(cudaMemcpyAsync(A.dev, A.host, bytes, cudaMemcpyHostToDevice, stream));
    gb::BatchTableGlobal T;
    T.images = (const gb::BatchImage*)A.dev;
    T.ctaStart = (const uint32_t*)((const uint8_t*)A.dev + bytesImages);
    rc = launch_batch_codec(codec, T, m, (uint32_t)total, stream);
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
