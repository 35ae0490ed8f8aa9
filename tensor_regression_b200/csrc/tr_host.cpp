// tr_host.cpp — host-side data movement of the C ABI: tr_upload copies a PAGEABLE host array into device
// memory at close to the pinned-memory DMA rate.
//
// cudaMemcpy from pageable memory stages through a small driver-owned pinned buffer on one thread.  Here the
// staging is explicit: a ring of pinned buffers, a pool of worker threads that fills a buffer in parallel slices
// (memcpy), and one asynchronous DMA per filled buffer on a private copy stream, so that the fill of buffer
// i+1 overlaps the DMA of buffer i.  A source that is already pinned / registered skips the staging.
// (The reference keeps X in host memory and calls .to(device) once, std:339-345 / mn:255: this is that copy.)
#include <cuda_runtime.h>
#include <immintrin.h>
#include <sched.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "tr_b200.h"

namespace {

// a tiny persistent worker pool: run(n, fn) executes fn(0..n-1) on the workers + the caller and returns when done
class Pool {
public:
    explicit Pool(int workers) {
        for (int i = 0; i < workers; ++i) th_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int workers() const { return (int)th_.size(); }
    void run(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; next_ = 0; total_ = n; done_ = 0; ++gen_;
        }
        cv_.notify_all();
        work();                                           // the caller takes slices too
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [this] { return done_ == total_; });
        fn_ = nullptr;
    }

private:
    void work() {
        for (;;) {
            int i;
            const std::function<void(int)>* f;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!fn_ || next_ >= total_) return;
                i = next_++;
                f = fn_;
            }
            (*f)(i);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (++done_ == total_) cv_done_.notify_all();
            }
        }
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, cv_done_;
    const std::function<void(int)>* fn_ = nullptr;
    int next_ = 0, total_ = 0, done_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
};

// Pageable -> pinned fill with NON-TEMPORAL stores: the staging buffer is written once and then read by the DMA
// engine only, so the stores bypass the cache and skip the read-for-ownership of every destination line (a plain
// memcpy of these 1 MB slices stays below glibc's non-temporal threshold and moves 3 bytes per byte copied).
// dst must be 32-byte aligned (staging buffers are page aligned, slices are 4 KB multiples).
__attribute__((target("avx2"))) static void copy_stream_avx2(char* dst, const char* src, size_t n) {
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i));
        const __m256i b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64));
        const __m256i d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
        _mm256_stream_si256((__m256i*)(dst + i), a);
        _mm256_stream_si256((__m256i*)(dst + i + 32), b);
        _mm256_stream_si256((__m256i*)(dst + i + 64), c);
        _mm256_stream_si256((__m256i*)(dst + i + 96), d);
    }
    _mm_sfence();
    if (i < n) memcpy(dst + i, src + i, n - i);
}

static void copy_slice(char* dst, const char* src, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("TR_B200_UPLOAD_PLAIN_MEMCPY");
    if (avx2 && ((uintptr_t)dst & 31) == 0) copy_stream_avx2(dst, src, n);
    else memcpy(dst, src, n);
}

struct Uploader {
    std::mutex mu;                       // one upload at a time per process (the staging ring is shared)
    int device = -1;
    size_t buf_bytes = 0;
    static constexpr int NB = 3;
    void* pin[NB] = {nullptr, nullptr, nullptr};
    cudaEvent_t done[NB] = {nullptr, nullptr, nullptr};
    cudaStream_t copy = nullptr;
    cudaEvent_t fence = nullptr;
    Pool* pool = nullptr;
    std::string err;
    double last_seconds = 0.0, last_fill_seconds = 0.0;
    int last_threads = 0, last_staged = 0;
};

Uploader& uploader() {
    static Uploader u;
    return u;
}

thread_local std::string g_host_error;

int hfail(int code, const std::string& msg) {
    g_host_error = msg;
    return code;
}

#define TRH_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            cudaGetLastError();                                                                     \
            return hfail(TR_ERR_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(e_));  \
        }                                                                                           \
    } while (0)

int ensure_ring(Uploader& u, int device, size_t buf_bytes, int threads) {
    if (u.device != device || u.buf_bytes < buf_bytes) {
        for (int i = 0; i < Uploader::NB; ++i) {
            if (u.pin[i]) { cudaFreeHost(u.pin[i]); u.pin[i] = nullptr; }
            if (u.done[i]) { cudaEventDestroy(u.done[i]); u.done[i] = nullptr; }
        }
        if (u.copy) { cudaStreamDestroy(u.copy); u.copy = nullptr; }
        if (u.fence) { cudaEventDestroy(u.fence); u.fence = nullptr; }
        u.device = device;
        u.buf_bytes = 0;
        for (int i = 0; i < Uploader::NB; ++i) {
            TRH_CUDA(cudaHostAlloc(&u.pin[i], buf_bytes, cudaHostAllocDefault));
            TRH_CUDA(cudaEventCreateWithFlags(&u.done[i], cudaEventDisableTiming));
        }
        TRH_CUDA(cudaStreamCreateWithFlags(&u.copy, cudaStreamNonBlocking));
        TRH_CUDA(cudaEventCreateWithFlags(&u.fence, cudaEventDisableTiming));
        u.buf_bytes = buf_bytes;
    }
    if (!u.pool || u.pool->workers() != threads - 1) {
        delete u.pool;
        u.pool = new Pool(std::max(0, threads - 1));
    }
    return TR_OK;
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

const char* tr_host_last_error(void) { return g_host_error.c_str(); }

int tr_upload(void* dst_device, const void* src_host, size_t bytes, int device, int threads, size_t chunk_bytes,
              void* stream) {
    if (!dst_device || !src_host) return hfail(TR_ERR_INVALID, "null pointer argument");
    if (bytes == 0) return TR_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return hfail(TR_ERR_CUDA, "no CUDA device available; this library has no CPU path");
    }
    if (device < 0 || device >= ndev) return hfail(TR_ERR_INVALID, "device out of range");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device);
    struct Restore { int p, d; ~Restore() { if (p >= 0 && p != d) cudaSetDevice(p); } } restore{prev, device};

    Uploader& u = uploader();
    std::lock_guard<std::mutex> lk(u.mu);
    if (threads <= 0) {
        unsigned hc = std::thread::hardware_concurrency();
        cpu_set_t cs;                                     // respect the affinity mask (ranks bound to a CPU subset)
        if (sched_getaffinity(0, sizeof(cs), &cs) == 0 && CPU_COUNT(&cs) > 0) hc = (unsigned)CPU_COUNT(&cs);
        threads = (int)std::min<unsigned>(hc ? hc : 4, 16);
        if (const char* ev = getenv("TR_B200_UPLOAD_THREADS")) { const int v = atoi(ev); if (v > 0) threads = v; }
    }
    if (chunk_bytes == 0) chunk_bytes = (size_t)32 << 20;
    chunk_bytes = std::min(std::max(chunk_bytes, (size_t)1 << 20), (size_t)1 << 30);
    chunk_bytes &= ~(size_t)4095;

    // already pinned / registered host memory: DMA straight from the source
    cudaPointerAttributes attr;
    bool pinned = false;
    if (cudaPointerGetAttributes(&attr, src_host) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
    else cudaGetLastError();

    const double t0 = now_s();
    cudaStream_t user = (cudaStream_t)stream;
    int rc = ensure_ring(u, device, pinned ? std::max<size_t>(u.buf_bytes, 4096) : chunk_bytes, pinned ? 1 : threads);
    if (rc) return rc;
    // the destination may still be in use by work queued on the caller's stream
    TRH_CUDA(cudaEventRecord(u.fence, user));
    TRH_CUDA(cudaStreamWaitEvent(u.copy, u.fence, 0));
    double fill_s = 0.0;
    if (pinned) {
        const size_t step = (size_t)1 << 30;
        for (size_t off = 0; off < bytes; off += step)
            TRH_CUDA(cudaMemcpyAsync((char*)dst_device + off, (const char*)src_host + off, std::min(step, bytes - off),
                                     cudaMemcpyHostToDevice, u.copy));
    } else {
        const char* src = (const char*)src_host;
        bool used[Uploader::NB] = {false, false, false};
        size_t off = 0;
        for (int i = 0; off < bytes; ++i) {
            const int s = i % Uploader::NB;
            const size_t len = std::min(chunk_bytes, bytes - off);
            if (used[s]) TRH_CUDA(cudaEventSynchronize(u.done[s]));      // the DMA that read this buffer is done
            // parallel fill: 4 slices per thread so that a slow core does not hold the chunk up
            const int nsl = std::max(1, std::min<int>(threads * 4, (int)((len + (1 << 20) - 1) >> 20)));
            const size_t sl = ((len + nsl - 1) / nsl + 4095) & ~(size_t)4095;
            char* dstp = (char*)u.pin[s];
            const char* srcp = src + off;
            const double f0 = now_s();
            u.pool->run(nsl, [&](int j) {
                const size_t a = (size_t)j * sl;
                if (a < len) copy_slice(dstp + a, srcp + a, std::min(sl, len - a));
            });
            fill_s += now_s() - f0;
            TRH_CUDA(cudaMemcpyAsync((char*)dst_device + off, u.pin[s], len, cudaMemcpyHostToDevice, u.copy));
            TRH_CUDA(cudaEventRecord(u.done[s], u.copy));
            used[s] = true;
            off += len;
        }
    }
    // the caller's stream continues after the last DMA; the host returns once the source may be modified
    TRH_CUDA(cudaEventRecord(u.fence, u.copy));
    TRH_CUDA(cudaStreamWaitEvent(user, u.fence, 0));
    TRH_CUDA(cudaEventSynchronize(u.fence));
    u.last_seconds = now_s() - t0;
    u.last_fill_seconds = fill_s;
    u.last_threads = pinned ? 0 : threads;
    u.last_staged = pinned ? 0 : 1;
    return TR_OK;
}

int tr_upload_stats(double* out4) {
    if (!out4) return TR_ERR_INVALID;
    Uploader& u = uploader();
    out4[0] = u.last_seconds; out4[1] = u.last_fill_seconds; out4[2] = (double)u.last_threads; out4[3] = (double)u.last_staged;
    return TR_OK;
}

}  // extern "C"
