// Host -> HBM staging for the blocking entry points (lvt_track & co.).
//
// The caller's images are pageable (lvt/src/lvt_c.cpp:69-70 borrows them for the call), so each one
// goes user memory -> pinned staging (same pitch as the pool) -> one contiguous
// DMA per band.  Done by a single thread that is ~45 us of memcpy per 1241x376 image before the first DMA can start, i.e. ~20 % of a
// frame.  Here the images are cut into row bands; a few lanes (the calling thread plus parked
// workers) pull bands off an atomic counter, copy the band to its place in the staging buffer and
// enqueue its DMA right away, so the copy engine runs while the other bands are still being staged.
// All bands are enqueued on the context's stream before stage() returns, hence anything launched on
// that stream afterwards sees the complete images.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include <mutex>
#include <thread>
#include <vector>

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
};

class UploadLanes
{
  public:
    static constexpr int kMaxBands = 32;

    UploadLanes() = default;
    UploadLanes(const UploadLanes &) = delete;
    UploadLanes &operator=(const UploadLanes &) = delete;

    void start(int device, int n_workers)
    {
        device_ = device;
        next_.store(1 << 30);
        for (int i = 0; i < n_workers; i++)
            workers_.emplace_back([this] { worker(); });
    }

    void stop()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
            generation_++;
        }
        cv_.notify_all();
        for (auto &t : workers_)
            t.join();
        workers_.clear();
    }

    ~UploadLanes() { stop(); }

    // cut [rows x width_bytes] into `parts` row bands and append them to the pending job
    void add_image(const void *src, size_t src_stride, void *stage, void *dst, size_t pitch, size_t width_bytes, int rows,
                   int parts)
    {
        const int step = (rows + parts - 1) / parts;
        for (int y = 0; y < rows && n_bands_ < kMaxBands; y += step)
        {
            UploadBand &b = bands_[n_bands_++];
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
        stream_ = stream;
        error_.store((int)cudaSuccess);
        done_.store(0);
        const int n = n_bands_;
        n_active_.store(n, std::memory_order_relaxed);
        next_.store(0, std::memory_order_release); // opens the job: bands_ / n_active_ are complete
        if (!workers_.empty() && n > 1)
        {
            {
                std::lock_guard<std::mutex> lk(mu_);
                generation_++;
            }
            cv_.notify_all();
        }
        lane();
        while (done_.load(std::memory_order_acquire) < n)
            ; // the other lanes are at most one band behind
        n_bands_ = 0;
        return (cudaError_t)error_.load();
    }

  private:
    void lane()
    {
        for (;;)
        {
            const int i = next_.fetch_add(1, std::memory_order_acq_rel);
            if (i >= n_active_.load(std::memory_order_relaxed) || i < 0)
                return;
            const UploadBand &b = bands_[i];
            // the staging buffer has the pool's pitch, so a band is one contiguous DMA (the padding
            // columns travel along; nothing reads them)
            if (b.src_stride == b.pitch)
                std::memcpy(b.stage, b.src, b.pitch * (size_t)(b.rows - 1) + b.width_bytes);
            else
                for (int y = 0; y < b.rows; y++)
                    std::memcpy(b.stage + (size_t)y * b.pitch, b.src + (size_t)y * b.src_stride, b.width_bytes);
            const cudaError_t e = cudaMemcpyAsync(b.dst, b.stage, b.pitch * (size_t)(b.rows - 1) + b.width_bytes,
                                                  cudaMemcpyHostToDevice, stream_);
            if (e != cudaSuccess)
                error_.store((int)e);
            done_.fetch_add(1, std::memory_order_release);
        }
    }

    void worker()
    {
        cudaSetDevice(device_);
        uint64_t seen = 0;
        for (;;)
        {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (quit_)
                    return;
            }
            lane();
        }
    }

    int device_ = 0;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    uint64_t generation_ = 0;
    bool quit_ = false;

    UploadBand bands_[kMaxBands];
    int n_bands_ = 0;
    std::atomic<int> n_active_{0}; // written before next_ is released
    cudaStream_t stream_ = nullptr;
    std::atomic<int> next_{1 << 30}, done_{0}, error_{0};
};

} // namespace lvtb
