// Host -> HBM staging for the blocking entry points (lvt_track & co.).
//
// The caller's images are pageable (lvt/src/lvt_c.cpp:69-70 borrows them for the call), so each one
// goes user memory -> pinned staging (same pitch as the pool) -> one contiguous DMA per row band.
// One thread needs ~50 us of memcpy per 1242x375 image before the first DMA can start, i.e. the
// longest host item on the way to the pose.  Here the image is cut into row bands; a few lanes (the
// calling thread plus helper threads) pull bands off an atomic counter and copy them into the
// staging buffer, while the calling thread alone talks to the driver: it enqueues the DMA of band b
// as soon as band b is staged (helpers never take the driver's locks, which the kernel launches
// that follow need).  Helpers spin for a short while after a job -- a caller that tracks frame
// after frame finds them awake -- and park on a condition variable when the stream of calls stops.
// All bands are enqueued on the given stream before run() returns, hence anything launched on that
// stream afterwards sees the complete images.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define LVT_CPU_RELAX() _mm_pause()
#else
#define LVT_CPU_RELAX() ((void)0)
#endif

namespace lvtb
{

struct UploadBand
{
    const uint8_t *src; // caller memory
    size_t src_stride;  // bytes
    uint8_t *stage;     // pinned, same pitch as the destination
    uint8_t *dst;       // device
    size_t pitch;       // of stage and dst
    size_t width_bytes;
    int rows;
    int dma_first; // first band of the DMA this band belongs to
    bool dma_last; // the DMA is enqueued once this band (and the ones before it) are staged
};

class UploadLanes
{
  public:
    static constexpr int kMaxBands = 32;
    static constexpr int kSpinMicros = 500; // helpers stay awake this long after a job

    UploadLanes() = default;
    UploadLanes(const UploadLanes &) = delete;
    UploadLanes &operator=(const UploadLanes &) = delete;

    void start(int device, int n_workers)
    {
        (void)device;
        for (int i = 0; i < n_workers; i++)
            workers_.emplace_back([this] { worker(); });
    }

    void stop()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_.store(true);
            generation_.fetch_add(1);
        }
        cv_.notify_all();
        for (auto &t : workers_)
            t.join();
        workers_.clear();
    }

    ~UploadLanes() { stop(); }

    int lanes() const { return 1 + (int)workers_.size(); }

    // cut [rows x width_bytes] into `parts` row bands and append them to the pending job
    // (the staging granularity) and `dmas` DMAs of consecutive bands (each DMA costs the calling thread
    // ~3 us of driver time, so there are fewer of them than bands)
    void add_image(const void *src, size_t src_stride, void *stage, void *dst, size_t pitch, size_t width_bytes, int rows,
                   int parts, int dmas = 1)
    {
        const int step = (rows + parts - 1) / parts;
        const int n_parts = (rows + step - 1) / step;
        dmas = dmas < 1 ? 1 : (dmas > n_parts ? n_parts : dmas);
        const int per_dma = (n_parts + dmas - 1) / dmas;
        int k = 0, first = n_bands_;
        for (int y = 0; y < rows && n_bands_ < kMaxBands; y += step, k++)
        {
            if (k % per_dma == 0)
                first = n_bands_;
            UploadBand &b = bands_[n_bands_++];
            b.dma_first = first;
            b.dma_last = (k % per_dma == per_dma - 1) || y + step >= rows || n_bands_ == kMaxBands;
            b.src = (const uint8_t *)src + (size_t)y * src_stride;
            b.src_stride = src_stride;
            b.stage = (uint8_t *)stage + (size_t)y * pitch;
            b.dst = (uint8_t *)dst + (size_t)y * pitch;
            b.pitch = pitch;
            b.width_bytes = width_bytes;
            b.rows = (y + step <= rows) ? step : rows - y;
        }
    }

    // stage + enqueue every pending band on `stream`; returns the first CUDA error (or cudaSuccess)
    cudaError_t run(cudaStream_t stream)
    {
        const int n = n_bands_;
        for (int i = 0; i < n; i++)
            staged_[i].store(0, std::memory_order_relaxed);
        // Jobs carry an epoch: a helper that drew its index from the previous job's (exhausted) counter
        // sees a job word of another epoch and drops it, so no band is ever copied twice.
        const uint64_t epoch = ++epoch_;
        job_.store(epoch << 32 | (uint32_t)n, std::memory_order_relaxed);
        next_.store(epoch << 32, std::memory_order_release); // opens the job: bands_ / job_ are complete
        if (!workers_.empty() && n > 1)
        {
            generation_.fetch_add(1);
            if (parked_.load() > 0)
            {
                std::lock_guard<std::mutex> lk(mu_);
                cv_.notify_all();
            }
        }
        cudaError_t err = cudaSuccess;
        int sent = 0;
        auto send_ready = [&](bool wait) {
            while (sent < n)
            {
                if (!staged_[sent].load(std::memory_order_acquire))
                {
                    if (!wait)
                        return;
                    LVT_CPU_RELAX();
                    continue; // a helper is at most one band behind
                }
                const UploadBand &b = bands_[sent];
                sent++;
                if (!b.dma_last)
                    continue;
                // the staging buffer has the pool's pitch, so consecutive bands are one contiguous DMA
                // (the padding columns travel along; nothing reads them)
                const UploadBand &f = bands_[b.dma_first];
                const size_t bytes = (size_t)(b.stage - f.stage) + b.pitch * (size_t)(b.rows - 1) + b.width_bytes;
                const cudaError_t e = cudaMemcpyAsync(f.dst, f.stage, bytes, cudaMemcpyHostToDevice, stream);
                if (e != cudaSuccess)
                    err = e;
            }
        };
        for (;;)
        {
            const int i = (int)(uint32_t)next_.fetch_add(1, std::memory_order_acq_rel); // the epoch is this job's
            if (i >= n)
                break;
            copy_band(i);
            send_ready(false);
        }
        send_ready(true);
        n_bands_ = 0;
        return err;
    }

  private:
    void copy_band(int i)
    {
        const UploadBand &b = bands_[i];
        if (b.src_stride == b.pitch)
            std::memcpy(b.stage, b.src, b.pitch * (size_t)(b.rows - 1) + b.width_bytes);
        else
            for (int y = 0; y < b.rows; y++)
                std::memcpy(b.stage + (size_t)y * b.pitch, b.src + (size_t)y * b.src_stride, b.width_bytes);
        staged_[i].store(1, std::memory_order_release);
    }

    void worker()
    {
        uint64_t seen = 0;
        auto last_job = std::chrono::steady_clock::now();
        for (;;)
        {
            if (generation_.load(std::memory_order_acquire) == seen)
            {
                if (std::chrono::steady_clock::now() - last_job < std::chrono::microseconds(kSpinMicros))
                {
                    LVT_CPU_RELAX();
                    continue; // spin: the next frame usually follows at once
                }
                std::unique_lock<std::mutex> lk(mu_);
                parked_.fetch_add(1);
                cv_.wait(lk, [&] { return generation_.load() != seen; });
                parked_.fetch_sub(1);
            }
            seen = generation_.load(std::memory_order_acquire);
            if (quit_.load())
                return;
            for (;;)
            {
                const uint64_t v = next_.fetch_add(1, std::memory_order_acq_rel);
                const uint64_t job = job_.load(std::memory_order_acquire);
                if ((v >> 32) != (job >> 32) || (uint32_t)v >= (uint32_t)job)
                    break; // exhausted, or an index of a job that is over
                copy_band((int)(uint32_t)v);
            }
            last_job = std::chrono::steady_clock::now();
        }
    }

    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::atomic<uint64_t> generation_{0};
    std::atomic<int> parked_{0};
    std::atomic<bool> quit_{false};

    UploadBand bands_[kMaxBands];
    int n_bands_ = 0;
    uint64_t epoch_ = 0;               // of the calling thread
    std::atomic<uint64_t> job_{0};     // epoch << 32 | bands of the job; written before next_ is released
    std::atomic<uint64_t> next_{~0ull >> 1}; // epoch << 32 | next band to hand out
    std::atomic<int> staged_[kMaxBands];
};

} // namespace lvtb
