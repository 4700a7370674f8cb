// host_copy.cpp — see host_copy.hpp.
#include "host_copy.hpp"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace mcskin {

void copy_rows_streaming(unsigned char* dst, const unsigned char* src, size_t dstPitch, size_t srcPitch, size_t rowBytes,
                         size_t rows) {
    for (size_t r = 0; r < rows; ++r) {
        unsigned char* d = dst + r * dstPitch;
        const unsigned char* s = src + r * srcPitch;
        size_t n = rowBytes;
#if defined(__SSE2__)
        // the destination is written once and not read here again: non-temporal stores spare the read-for-ownership
        if (n >= 256 && ((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(s)) & 15u) == 0) {
            const size_t blocks = n / 64;
            for (size_t b = 0; b < blocks; ++b) {
                const __m128i v0 = _mm_load_si128(reinterpret_cast<const __m128i*>(s) + 0);
                const __m128i v1 = _mm_load_si128(reinterpret_cast<const __m128i*>(s) + 1);
                const __m128i v2 = _mm_load_si128(reinterpret_cast<const __m128i*>(s) + 2);
                const __m128i v3 = _mm_load_si128(reinterpret_cast<const __m128i*>(s) + 3);
                _mm_stream_si128(reinterpret_cast<__m128i*>(d) + 0, v0);
                _mm_stream_si128(reinterpret_cast<__m128i*>(d) + 1, v1);
                _mm_stream_si128(reinterpret_cast<__m128i*>(d) + 2, v2);
                _mm_stream_si128(reinterpret_cast<__m128i*>(d) + 3, v3);
                s += 64;
                d += 64;
            }
            n -= blocks * 64;
        }
#endif
        if (n) std::memcpy(d, s, n);
    }
#if defined(__SSE2__)
    _mm_sfence();
#endif
}

namespace {

class CopyPool {
public:
    static CopyPool& instance() {
        static CopyPool pool;
        return pool;
    }

    bool run(int device, const std::vector<HostCopyJob>& jobs) {
        if (jobs.empty()) return true;
        std::lock_guard<std::mutex> one(callMu_);  // one set of jobs at a time
        start_threads();
        {
            std::lock_guard<std::mutex> lock(mu_);
            jobs_ = &jobs;
            device_ = device;
            next_.store(0);
            pending_ = jobs.size();
            failed_.store(false);
            ++generation_;
        }
        wake_.notify_all();
        work(device);  // the caller copies too
        std::unique_lock<std::mutex> lock(mu_);
        done_.wait(lock, [&] { return pending_ == 0; });
        jobs_ = nullptr;
        return !failed_.load();
    }

private:
    CopyPool() = default;
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            stop_ = true;
        }
        wake_.notify_all();
        for (std::thread& t : threads_) t.join();
    }

    void start_threads() {
        if (!threads_.empty()) return;
        const unsigned int hw = std::max(1u, std::thread::hardware_concurrency());
        // B200 box, 16 cores, 33 MB frame: 1 thread 3.5 ms per frame, 3 or 5 threads 1.70, 8 threads 1.36, 12 threads 1.56
        // (the host's memory carries the frame three times: DMA in, read, write)
        unsigned int n = hw >= 4 ? std::min(7u, hw / 2 - 1) : 1u;
        if (const char* v = std::getenv("MCSKIN_COPY_THREADS")) n = static_cast<unsigned int>(std::min(64, std::max(0, std::atoi(v))));
        for (unsigned int i = 0; i < n; ++i) threads_.emplace_back([this] { worker(); });
    }

    void worker() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lock(mu_);
                wake_.wait(lock, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
            }
            work(0);
        }
    }

    // Takes jobs until none is left.  (jobs_ stays valid until pending_ reaches 0, and a thread that finds no job
    // never touches it.)
    void work(int) {
        int deviceSet = -1;
        for (;;) {
            const std::vector<HostCopyJob>* jobs;
            size_t i;
            int device;
            {
                std::lock_guard<std::mutex> lock(mu_);
                jobs = jobs_;
                if (!jobs) return;
                i = next_.load();
                if (i >= jobs->size()) return;
                next_.store(i + 1);
                device = device_;  // (a thread that woke late may find the jobs of the next call)
            }
            const HostCopyJob& j = (*jobs)[i];
            if (j.after) {
                if (deviceSet != device) {
                    cudaSetDevice(device);
                    deviceSet = device;
                }
                if (cudaEventSynchronize(j.after) != cudaSuccess) failed_.store(true);
            }
            if (!failed_.load()) copy_rows_streaming(j.dst, j.src, j.dstPitch, j.srcPitch, j.rowBytes, j.rows);
            bool last;
            {
                std::lock_guard<std::mutex> lock(mu_);
                last = --pending_ == 0;
            }
            if (last) done_.notify_all();
        }
    }

    std::mutex callMu_, mu_;
    std::condition_variable wake_, done_;
    std::vector<std::thread> threads_;
    const std::vector<HostCopyJob>* jobs_ = nullptr;
    std::atomic<size_t> next_{0};
    std::atomic<bool> failed_{false};
    size_t pending_ = 0;
    unsigned long long generation_ = 0;
    int device_ = 0;
    bool stop_ = false;
};

}  // namespace

bool run_host_copies(int device, const std::vector<HostCopyJob>& jobs) { return CopyPool::instance().run(device, jobs); }

}  // namespace mcskin
