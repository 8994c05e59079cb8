// Error reporting / bookkeeping entry points of libscrib200.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace scrib200 {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace scrib200

extern "C" {

int scrib200_version(void) { return 100; }

const char* scrib200_last_error(void) { return scrib200::g_error; }

int64_t scrib200_launch_count(void) { return scrib200::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Host -> device staging for pageable host arrays (numpy memory).  A pageable cudaMemcpy runs at ~10 GB/s and
// framework-level parallel copies fight with whatever thread pools the host process already spins (BLAS, OpenMP), so
// the library owns the path: a ring of pinned staging buffers, a few plain memcpy worker threads filling chunk i+1
// while the copy engine drains chunk i.  On return every byte of `src` has been read (the caller may reuse it); the
// last DMAs are ordered on `stream` like any other asynchronous copy.
#include <stdlib.h>

#include <condition_variable>
#include <mutex>
#include <emmintrin.h>

#include <thread>
#include <vector>

namespace scrib200 {

// Pageable -> pinned staging copy with non-temporal stores: the destination is written once and read next by the DMA
// engine, so it should not be pulled into the cache first (no read-for-ownership traffic: ~1.4x a plain memcpy on the
// hosts measured).  `d` is 16-byte aligned (slices of the page-aligned staging buffers start at multiples of 4096).
static void stream_copy(char* d, const char* s, size_t n) {
    size_t i = 0;
    if ((reinterpret_cast<uintptr_t>(d) & 15u) == 0) {
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
            const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
        }
        _mm_sfence();
    }
    if (i < n) memcpy(d + i, s + i, n - i);
}

class CopyPool {
  public:
    explicit CopyPool(int n) : pending_(0), stop_(false) {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // copy [src, src+n) to dst split into `parts` slices; returns when all slices are done
    void copy(char* dst, const char* src, size_t n, int parts) {
        const size_t step = ((n / parts) + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> lk(m_);
            for (size_t off = step; off < n; off += step) {
                jobs_.push_back({dst + off, src + off, (n - off < step) ? n - off : step});
                ++pending_;
            }
        }
        cv_.notify_all();
        stream_copy(dst, src, n < step ? n : step);   // the caller's share
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

  private:
    struct Job { char* d; const char* s; size_t n; };
    void loop() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this] { return stop_ || !jobs_.empty(); });
                if (stop_ && jobs_.empty()) return;
                j = jobs_.back();
                jobs_.pop_back();
            }
            stream_copy(j.d, j.s, j.n);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::vector<Job> jobs_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    int pending_;
    bool stop_;
};

constexpr int STAGE_BUFS = 4;
constexpr size_t STAGE_BYTES = (size_t)8 << 20;

struct Stager {
    char* buf[STAGE_BUFS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[STAGE_BUFS];
    bool used[STAGE_BUFS] = {false, false, false, false};
    int device = -1;
    CopyPool* pool = nullptr;
    int parts = 1;
    std::mutex m;
};
static Stager g_stager;

}  // namespace scrib200

extern "C" int scrib200_h2d(void* dst_device, const void* src_host, size_t nbytes, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(dst_device && src_host, "h2d: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, src_host) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
        // already page-locked: one DMA, no staging
        cudaError_t e = cudaMemcpyAsync(dst_device, src_host, nbytes, cudaMemcpyHostToDevice, st);
        SCRIB200_REQUIRE(e == cudaSuccess, "h2d: %s", cudaGetErrorString(e));
        return SCRIB200_OK;
    }
    cudaGetLastError();   // clear the "not a CUDA pointer" status older drivers set
    std::lock_guard<std::mutex> lk(g_stager.m);
    Stager& S = g_stager;
    int dev = 0;
    cudaGetDevice(&dev);
    if (S.buf[0] == nullptr) {
        for (int i = 0; i < STAGE_BUFS; ++i) {
            cudaError_t e = cudaHostAlloc((void**)&S.buf[i], STAGE_BYTES, cudaHostAllocPortable);
            SCRIB200_REQUIRE(e == cudaSuccess, "h2d: cudaHostAlloc failed: %s", cudaGetErrorString(e));
        }
        int workers = (int)std::thread::hardware_concurrency() - 1;
        if (const char* env = getenv("SCRIB200_COPY_THREADS")) workers = atoi(env) - 1;
        workers = workers < 0 ? 0 : (workers > 15 ? 15 : workers);
        S.pool = new CopyPool(workers);
        S.parts = workers + 1;
    }
    if (S.device != dev) {   // events belong to a device
        for (int i = 0; i < STAGE_BUFS; ++i) {
            if (S.device >= 0) cudaEventDestroy(S.ev[i]);
            cudaEventCreateWithFlags(&S.ev[i], cudaEventDisableTiming);
            S.used[i] = false;
        }
        S.device = dev;
    }
    const char* src = reinterpret_cast<const char*>(src_host);
    char* dst = reinterpret_cast<char*>(dst_device);
    int b = 0;
    for (size_t off = 0; off < nbytes; off += STAGE_BYTES, b = (b + 1) % STAGE_BUFS) {
        const size_t n = (nbytes - off < STAGE_BYTES) ? nbytes - off : STAGE_BYTES;
        if (S.used[b]) cudaEventSynchronize(S.ev[b]);   // the DMA that last read this staging buffer is done
        S.pool->copy(S.buf[b], src + off, n, S.parts);
        cudaError_t e = cudaMemcpyAsync(dst + off, S.buf[b], n, cudaMemcpyHostToDevice, st);
        SCRIB200_REQUIRE(e == cudaSuccess, "h2d: %s", cudaGetErrorString(e));
        cudaEventRecord(S.ev[b], st);
        S.used[b] = true;
    }
    return SCRIB200_OK;
}
