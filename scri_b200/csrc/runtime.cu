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
// while the copy engine drains chunk i.  For pageable `src` every byte has been read on return (the caller may reuse
// it) and the last DMAs are ordered on `stream`.  Page-locked `src` (cudaHostAlloc'd, or registered through
// scrib200_host_register) is copied by a single asynchronous DMA instead: it must stay unchanged until `stream` has
// passed this point, as for any cudaMemcpyAsync.
#include <stdlib.h>

#include <condition_variable>
#include <mutex>
#include <emmintrin.h>
#include <sched.h>

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
        // copy threads: this process's share of the cores it may run on (its CPU affinity mask divided by the ranks of
        // the node, LOCAL_WORLD_SIZE under torchrun) - eight ranks each spawning one thread per core of the whole host
        // would oversubscribe it 8x; the calling thread is one of the copiers
        int cores = (int)std::thread::hardware_concurrency();
        cpu_set_t mask;
        if (sched_getaffinity(0, sizeof(mask), &mask) == 0 && CPU_COUNT(&mask) > 0) cores = CPU_COUNT(&mask);
        int ranks = 1;
        if (const char* env = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(env) > 0 ? atoi(env) : 1;
        int workers = cores / ranks - 1;
        if (const char* env = getenv("SCRIB200_COPY_THREADS")) workers = atoi(env) - 1;
        workers = workers < 0 ? 0 : (workers > 15 ? 15 : workers);
        S.pool = new CopyPool(workers);
        S.parts = workers + 1;
    }
    if (S.device != dev) {   // events belong to a device
        for (int i = 0; i < STAGE_BUFS; ++i) {
            if (S.device >= 0) {
                if (S.used[i]) cudaEventSynchronize(S.ev[i]);   // a DMA of the previous device may still read this buffer
                cudaEventDestroy(S.ev[i]);
            }
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


// Page-lock a caller's array in place so that scrib200_h2d DMAs straight from it: no staging copy, no CPU work, and
// none of the three host-DRAM crossings of the staged path.  Worth it for arrays that go up more than once (the frame-
// fixing optimisers transform one waveform by many transformations); registering costs about as much as one staged copy.
extern "C" int scrib200_host_register(const void* p, size_t nbytes) {
    using namespace scrib200;
    SCRIB200_REQUIRE(p && nbytes > 0, "host_register: null pointer");
    // memory that is page-locked already (cudaHostAlloc'd, e.g. a numpy view of a pinned torch tensor, or registered by
    // someone else) needs nothing - and cudaHostRegister on it fails with "invalid argument"
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess) {
        if (attr.type == cudaMemoryTypeHost) return 1;      // page-locked by someone else: theirs to release
    } else {
        cudaGetLastError();
    }
    cudaError_t e = cudaHostRegister(const_cast<void*>(p), nbytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) cudaGetLastError();      // never leave the error for the next launch check to find
    if (e == cudaErrorHostMemoryAlreadyRegistered) return 1;
    SCRIB200_REQUIRE(e == cudaSuccess, "host_register: %s", cudaGetErrorString(e));
    return SCRIB200_OK;
}

extern "C" int scrib200_host_unregister(const void* p) {
    using namespace scrib200;
    SCRIB200_REQUIRE(p, "host_unregister: null pointer");
    cudaError_t e = cudaHostUnregister(const_cast<void*>(p));
    if (e != cudaSuccess) cudaGetLastError();      // not registered (any more): nothing to undo
    return SCRIB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Host-side rotor integration dR/dt = omega(t) R / 2 (quaternion.integrate_angular_velocity as scri/mode_calculations.py:467
// calls it).  The problem is one serial ODE over the whole series - nothing for a GPU - but the reference's route (scipy's
// DOP853 with a Python right-hand side: ~60 000 interpreter round trips for 2e4 samples, 1.3 s) is what dominates
// to_corotating_frame once the mode kernels run on the device, so the integrator lives here as native code: Dormand-Prince
// 8(5,3) with the standard step controller, steps clipped to the sample times (the output needs no dense interpolant),
// omega = the not-a-knot cubic spline through the samples, evaluated from its piecewise-polynomial coefficients.
namespace scrib200 {
#include "dop853_tables.inc"

struct RotorRhs {
    const double* t;      // [n]
    const double* coef;   // [4][n-1][3]: scipy PPoly layout, coef[k][i][c] multiplies (x - t_i)^(3-k)
    int64_t n;
    inline void operator()(int64_t i, double x, const double* y, double* f) const {
        const double dx = x - t[i];
        const int64_t s = (n - 1) * 3;
        const double* c0 = coef + i * 3;
        double w[3];
        for (int c = 0; c < 3; ++c) w[c] = ((c0[c] * dx + c0[s + c]) * dx + c0[2 * s + c]) * dx + c0[3 * s + c];
        // (0, w) * y / 2
        f[0] = 0.5 * (-w[0] * y[1] - w[1] * y[2] - w[2] * y[3]);
        f[1] = 0.5 * (w[0] * y[0] + w[1] * y[3] - w[2] * y[2]);
        f[2] = 0.5 * (-w[0] * y[3] + w[1] * y[0] + w[2] * y[1]);
        f[3] = 0.5 * (w[0] * y[2] - w[1] * y[1] + w[2] * y[0]);
    }
};

}  // namespace scrib200

extern "C" int scrib200_integrate_angular_velocity(const double* t, int64_t n, const double* coef, const double* R0, double atol,
                                                   double rtol, double* out, int64_t* n_rhs) {
    using namespace scrib200;
    SCRIB200_REQUIRE(t && coef && R0 && out, "integrate_angular_velocity: null pointer");
    SCRIB200_REQUIRE(n >= 2, "integrate_angular_velocity: needs at least two samples; got %lld", (long long)n);
    SCRIB200_REQUIRE(atol > 0.0 || rtol > 0.0, "integrate_angular_velocity: atol and rtol cannot both be zero");
    const RotorRhs rhs{t, coef, n};
    constexpr int NS = 12, D = 4;
    const double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0, EXPO = -1.0 / 8.0;
    double y[D], K[NS + 1][D], ynew[D];
    for (int c = 0; c < D; ++c) out[c] = y[c] = R0[c];
    int64_t evals = 1;
    rhs(0, t[0], y, K[0]);
    double h_next = t[1] - t[0];
    for (int64_t i = 0; i + 1 < n; ++i) {
        const double t_end = t[i + 1];
        SCRIB200_REQUIRE(t_end > t[i], "integrate_angular_velocity: times must be strictly increasing (sample %lld)", (long long)i);
        double tc = t[i];
        while (tc < t_end) {
            double h = h_next;
            bool clipped = false;
            if (tc + h >= t_end || t_end - (tc + h) < 1e-14 * fabs(t_end)) {
                h = t_end - tc;
                clipped = true;
            }
            bool accepted = false;
            while (!accepted) {
                SCRIB200_REQUIRE(fabs(h) > 16.0 * 2.220446049250313e-16 * fmax(fabs(tc), 1.0), "integrate_angular_velocity: step size underflow at t=%g", tc);
                for (int s = 1; s < NS; ++s) {
                    double ys[D];
                    for (int c = 0; c < D; ++c) {
                        double acc = 0.0;
                        for (int j = 0; j < s; ++j) acc += DOP_A[s][j] * K[j][c];
                        ys[c] = y[c] + h * acc;
                    }
                    rhs(i, tc + DOP_C[s] * h, ys, K[s]);
                }
                for (int c = 0; c < D; ++c) {
                    double acc = 0.0;
                    for (int j = 0; j < NS; ++j) acc += DOP_B[j] * K[j][c];
                    ynew[c] = y[c] + h * acc;
                }
                rhs(i, tc + h, ynew, K[NS]);
                evals += NS;
                double e5 = 0.0, e3 = 0.0;
                for (int c = 0; c < D; ++c) {
                    const double scale = atol + fmax(fabs(y[c]), fabs(ynew[c])) * rtol;
                    double a5 = 0.0, a3 = 0.0;
                    for (int j = 0; j <= NS; ++j) {
                        a5 += DOP_E5[j] * K[j][c];
                        a3 += DOP_E3[j] * K[j][c];
                    }
                    e5 += (a5 / scale) * (a5 / scale);
                    e3 += (a3 / scale) * (a3 / scale);
                }
                double err = 0.0;
                if (e5 != 0.0 || e3 != 0.0) err = fabs(h) * e5 / sqrt((e5 + 0.01 * e3) * D);
                if (err < 1.0) {
                    const double factor = (err == 0.0) ? MAX_FACTOR : fmin(MAX_FACTOR, SAFETY * pow(err, EXPO));
                    if (!clipped || factor < 1.0) h_next = h * factor;          // a clipped step says nothing about growing
                    else h_next = fmax(h_next, h * factor);
                    accepted = true;
                } else {
                    h *= fmax(MIN_FACTOR, SAFETY * pow(err, EXPO));
                    clipped = false;
                }
            }
            tc = clipped ? t_end : tc + h;
            for (int c = 0; c < D; ++c) {
                y[c] = ynew[c];
                K[0][c] = K[NS][c];          // first same as last
            }
        }
        for (int c = 0; c < D; ++c) out[(i + 1) * D + c] = y[c];
        if (i + 2 < n) rhs(i + 1, t_end, y, K[0]);                      // same value, evaluated in the next interval's polynomial
    }
    if (n_rhs) *n_rhs = evals;
    return SCRIB200_OK;
}
