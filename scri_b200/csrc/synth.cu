// K1: SWSH synthesis as one real FP64 tensor-core GEMM with the BMS epilogue fused.
//
// Replaces scri/waveform_grid.py:475-484 (np.tensordot -> zgemm), :486-503 (constant data-type
// correction) and :559 (conformal factor):   F[i,g] = (sum_lm a[i,lm] Y[g,lm] - c[g]) * k[g]^w.
//
// The complex product is folded into a real GEMM: A = modes viewed as real [N, K=2n] (re,im
// interleaved, exactly the caller's memory), B = packed real table [Kpad, Ncpad] with
//   B[2lm, 2g] = Re Y, B[2lm+1, 2g] = -Im Y, B[2lm, 2g+1] = Im Y, B[2lm+1, 2g+1] = Re Y
// so that C = A.B is F with re/im interleaved.  8*n*G flops per time step, none redundant.
//
// sm_100a has no tcgen05 kind for f64; the FP64 tensor path is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).
// CTA tile 64x64, 4 warps as 2(M) x 2(N), warp tile 32x32 = 4x4 DMMA tiles (32 accumulator doubles per thread), four
// CTAs per SM (small independent CTAs beat 128x64 ones by 12 %: fewer warps wait on each barrier); K is streamed in
// BK=8 slabs through a 4-stage cp.async ring; shared-memory strides BK+4 and 68 doubles for the fragment loads.
#include <stdlib.h>

#include "common.cuh"

namespace scrib200 {

constexpr int BN = 64;
constexpr int BS = BN + 4;   // B smem row stride (doubles)
template <int BK, int STAGES, int BM = 128>
struct SynthCfg {
    static constexpr int THREADS = 2 * BM;   // (BM / 32) x 2 warps of 32 x 32
    static constexpr int AS = BK + 4;   // A smem row stride (doubles)
    static constexpr int A_STAGE = BM * AS;
    static constexpr int B_STAGE = BK * BS;
    static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int BK, int STAGES, int BM>
__global__ void __launch_bounds__(2 * BM, 256 / BM)
swsh_synth_dmma_kernel(const double* __restrict__ A, int64_t M, int K, const double* __restrict__ B, int Kpad,
                       int Ncpad, const double* __restrict__ offset, const double* __restrict__ scale, int Nc,
                       double* __restrict__ C, int band) {
    constexpr int AS = SynthCfg<BK, STAGES, BM>::AS, A_STAGE = SynthCfg<BK, STAGES, BM>::A_STAGE, B_STAGE = SynthCfg<BK, STAGES, BM>::B_STAGE;
    constexpr int SYNTH_THREADS = 2 * BM;
    extern __shared__ __align__(16) double smem_d[];
    double* sA = smem_d;
    double* sB = smem_d + STAGES * A_STAGE;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;           // 4 x 2 warps
    // Tile walk: CTAs are dealt in launch order to bands of `band` column tiles; inside a band the column tiles of one row
    // tile come first (they share the A row tile through L2), then the next row tile.  The B panel of a band (band x Kpad
    // x 64 doubles, sized by the host to sit in L2) is then read from HBM once, and A once per band - without this a
    // table larger than L2 (ell_max = 32: 616 MB) was streamed from HBM once per row tile.
    int bx = blockIdx.x, by = blockIdx.y;
    if (band < (int)gridDim.x) {
        const int64_t lin = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
        const int64_t per_band = (int64_t)band * gridDim.y;
        const int bnd = (int)(lin / per_band);
        const int width = min(band, (int)gridDim.x - bnd * band);      // the last band may be narrower
        const int64_t rem = lin - (int64_t)bnd * per_band;
        by = (int)(rem / width);
        bx = bnd * band + (int)(rem - (int64_t)by * width);
    }
    const int64_t row0 = (int64_t)by * BM;
    const int col0 = bx * BN;
    const int KT = Kpad / BK;

    auto load_stage = [&](int stage, int kt) {
        const int k0 = kt * BK;
        // A: 128 rows x 16 doubles = 128 x 8 chunks of 16 B -> 4 chunks per thread
        double* a_dst = sA + stage * A_STAGE;
#pragma unroll
        for (int c = 0; c < (BM * BK / 2) / SYNTH_THREADS; ++c) {
            int chunk = tid + c * SYNTH_THREADS;
            int r = chunk / (BK / 2), kc = (chunk % (BK / 2)) * 2;
            int64_t gr = row0 + r;
            int k = k0 + kc;
            int bytes = (gr < M && k < K) ? 16 : 0;   // K is even, so a chunk is all-in or all-out
            const double* src = A + (bytes ? (gr * K + k) : 0);
            cp_async16(a_dst + r * AS + kc, src, bytes);
        }
        // B: 16 rows x 64 doubles = 16 x 32 chunks -> 2 chunks per thread (table is pre-padded)
        double* b_dst = sB + stage * B_STAGE;
#pragma unroll
        for (int c = 0; c < (BK * BN / 2) / SYNTH_THREADS; ++c) {
            int chunk = tid + c * SYNTH_THREADS;
            int r = chunk >> 5, nc = (chunk & 31) * 2;
            cp_async16(b_dst + r * BS + nc, B + (size_t)(k0 + r) * Ncpad + col0 + nc, 16);
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    const int ar = lane >> 2, ak = lane & 3;   // A fragment: row = lane/4, k = lane%4
    const int bk = lane & 3, bn = lane >> 2;   // B fragment: k = lane%4, n = lane/4

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const double* a_s = sA + (kt % STAGES) * A_STAGE + (wm * 32 + ar) * AS + ak;
        const double* b_s = sB + (kt % STAGES) * B_STAGE + bk * BS + wn * 32 + bn;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = a_s[i * 8 * AS + ks * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = b_s[ks * 4 * BS + j * 8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: (acc - offset[col]) * scale[col]; C fragment: row = lane/4, cols = 2*(lane%4) + {0,1}
    const int cr = lane >> 2, cc = (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int col = col0 + wn * 32 + j * 8 + cc;
        if (col >= Nc) continue;
        double2 off = *reinterpret_cast<const double2*>(offset + col);
        double2 scl = *reinterpret_cast<const double2*>(scale + col);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int64_t row = row0 + wm * 32 + i * 8 + cr;
            if (row >= M) continue;
            double2 v;
            v.x = (acc[i][j][0] - off.x) * scl.x;
            v.y = (acc[i][j][1] - off.y) * scl.y;
            *reinterpret_cast<double2*>(C + row * Nc + col) = v;
        }
    }
}

}  // namespace scrib200

extern "C" int scrib200_swsh_synthesize(const double* modes, int64_t n_times, int n_modes, const double* Bmat,
                                        int Kpad, int Ncpad, const double* offset, const double* scale, int G,
                                        double* F, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(modes && Bmat && offset && scale && F, "swsh_synthesize: null pointer");
    SCRIB200_REQUIRE(n_modes > 0 && G > 0, "swsh_synthesize: bad sizes n_modes=%d G=%d", n_modes, G);
    const int K = 2 * n_modes, Nc = 2 * G;
    SCRIB200_REQUIRE(Kpad % 16 == 0 && Kpad >= K, "swsh_synthesize: Kpad=%d must be a multiple of 16 >= %d", Kpad, K);
    SCRIB200_REQUIRE(Ncpad % BN == 0 && Ncpad >= Nc, "swsh_synthesize: Ncpad=%d must be a multiple of %d >= %d", Ncpad,
                     BN, Nc);
    SCRIB200_REQUIRE(aligned16(modes) && aligned16(Bmat) && aligned16(F) && aligned16(offset) && aligned16(scale),
                     "swsh_synthesize: pointers must be 16-byte aligned");
    if (n_times <= 0) return SCRIB200_OK;
    int variant = 0;
    if (const char* env = getenv("SCRIB200_SYNTH_VARIANT")) variant = atoi(env);
    // column tiles fastest so the CTAs sharing an A row-tile run together (A is then read from HBM once);
    // gridDim.y is limited to 65535, so very long series go in slabs of rows
    const int64_t max_rows = (int64_t)65535 * 64;
    // column tiles per band: 48 MB of the table (B200: 126 MB of L2, shared with the A rows and the output in flight)
    int band = (int)(((size_t)48 << 20) / ((size_t)Kpad * BN * sizeof(double)));
    if (band < 1) band = 1;
    for (int64_t r0 = 0; r0 < n_times; r0 += max_rows) {
        int64_t rows = n_times - r0 < max_rows ? n_times - r0 : max_rows;
#define SYNTH_LAUNCH(BK_, ST_, BM_)                                                                                    \
    do {                                                                                                               \
        dim3 grid_(Ncpad / BN, (unsigned)((rows + BM_ - 1) / BM_));                                                    \
        cudaFuncSetAttribute(swsh_synth_dmma_kernel<BK_, ST_, BM_>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                             (int)SynthCfg<BK_, ST_, BM_>::SMEM);                                                      \
        swsh_synth_dmma_kernel<BK_, ST_, BM_><<<grid_, 2 * BM_, SynthCfg<BK_, ST_, BM_>::SMEM, (cudaStream_t)stream>>>( \
            modes + r0 * K, rows, K, Bmat, Kpad, Ncpad, offset, scale, Nc, F + r0 * Nc, band);                         \
    } while (0)
        // measured at config 2 (K = 160): 128x64 / BK16 / 3 stages 1.47 ms, 64x64 / BK16 / 3 stages 1.33 ms, 64x64 / BK8 / 4 stages 1.30 ms
        if (variant == 1) SYNTH_LAUNCH(16, 3, 128);
        else if (variant == 2) SYNTH_LAUNCH(16, 3, 64);
        else SYNTH_LAUNCH(8, 4, 64);
#undef SYNTH_LAUNCH
        SCRIB200_CHECK_LAUNCH("swsh_synthesize");
    }
    return SCRIB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// K1 by the three-multiplication ("3M") complex product (round 2).  F = a Y needs, per complex term,
//     T1 = a_r Y_r,   T2 = a_i Y_i,   T3 = (a_r + a_i)(Y_r + Y_i):     Re = T1 - T2,   Im = T3 - T1 - T2,
// i.e. three real GEMMs [N, n] x [n, G] - 6 n G flops instead of the 8 n G of the folded real GEMM above, on a kernel that
// sits at 86 % DMMA-pipe activity.  The A operand is the caller's complex array: one 16-byte shared-memory load gives a lane
// (a_r, a_i) of its fragment element and a_r + a_i costs one add; the table comes as three real planes Y_r, Y_i, Y_r + Y_i
// (scrib200_swsh_pack3m, from the packed table of the folded kernel).  The sums carry the rounding of a_r + a_i and
// Y_r + Y_i, a few ulp of |a| |Y| per term - the same size as the rounding of the products themselves.
// CTA tile 64 rows x 32 complex columns, 4 warps as 2 x 2, warp tile 32 x 16 complex = 4 x 2 DMMA tiles x 3 products
// (48 accumulator doubles per thread), K streamed in slabs of 8 modes through a 3-stage cp.async ring, three CTAs per SM.
namespace scrib200 {

constexpr int M3_BM = 64, M3_BNC = 32, M3_BKM = 8;
constexpr int M3_AS = M3_BKM + 4;        // A row stride in double2 (12: two rows of a quarter-warp hit different bank halves)
constexpr int M3_BS = M3_BNC + 4;        // B row stride in doubles
constexpr int M3_A_STAGE = M3_BM * M3_AS;            // double2
constexpr int M3_B_STAGE = 3 * M3_BKM * M3_BS;       // double
template <int STAGES>
constexpr size_t m3_smem() { return (size_t)STAGES * (M3_A_STAGE * sizeof(double2) + M3_B_STAGE * sizeof(double)); }

__global__ void __launch_bounds__(256)
swsh_pack3m_kernel(const double* __restrict__ Bmat, int Ncpad, int n_modes, int G, double* __restrict__ B3, int npad, int Gpad) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)npad * Gpad) return;
    const int lm = (int)(idx / Gpad), g = (int)(idx - (int64_t)lm * Gpad);
    double yr = 0.0, yi = 0.0;
    if (lm < n_modes && g < G) {
        yr = Bmat[(size_t)(2 * lm) * Ncpad + 2 * g];
        yi = Bmat[(size_t)(2 * lm) * Ncpad + 2 * g + 1];
    }
    const size_t plane = (size_t)npad * Gpad;
    B3[idx] = yr;
    B3[plane + idx] = yi;
    B3[2 * plane + idx] = yr + yi;
}

template <int M3_STAGES>
__global__ void __launch_bounds__(128, M3_STAGES <= 3 ? 3 : 2)
swsh_synth3m_kernel(const double2* __restrict__ A, int64_t M, int n, const double* __restrict__ B3, int npad, int Gpad,
                    const double* __restrict__ offset, const double* __restrict__ scale, int G, double2* __restrict__ C, int band) {
    extern __shared__ __align__(16) unsigned char smem3[];
    double2* sA = reinterpret_cast<double2*>(smem3);
    double* sB = reinterpret_cast<double*>(sA + M3_STAGES * M3_A_STAGE);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;             // 2 x 2 warps
    int bx = blockIdx.x, by = blockIdx.y;
    if (band < (int)gridDim.x) {                          // band-wise tile walk, as in swsh_synth_dmma_kernel
        const int64_t lin = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
        const int64_t per_band = (int64_t)band * gridDim.y;
        const int bnd = (int)(lin / per_band);
        const int width = min(band, (int)gridDim.x - bnd * band);
        const int64_t rem = lin - (int64_t)bnd * per_band;
        by = (int)(rem / width);
        bx = bnd * band + (int)(rem - (int64_t)by * width);
    }
    const int64_t row0 = (int64_t)by * M3_BM;
    const int col0 = bx * M3_BNC;                         // first complex column
    const int KT = npad / M3_BKM;
    const size_t plane = (size_t)npad * Gpad;

    auto load_stage = [&](int stage, int kt) {
        const int k0 = kt * M3_BKM;
        double2* a_dst = sA + stage * M3_A_STAGE;
#pragma unroll
        for (int c = 0; c < (M3_BM * M3_BKM) / 128; ++c) {       // 64 rows x 8 modes of 16 bytes: 4 chunks per thread
            const int chunk = tid + c * 128;
            const int r = chunk / M3_BKM, kc = chunk % M3_BKM;
            const int64_t gr = row0 + r;
            const int k = k0 + kc;
            const int bytes = (gr < M && k < n) ? 16 : 0;
            cp_async16(a_dst + r * M3_AS + kc, A + (bytes ? (gr * n + k) : 0), bytes);
        }
        double* b_dst = sB + stage * M3_B_STAGE;
#pragma unroll
        for (int c = 0; c < (3 * M3_BKM * M3_BNC / 2) / 128; ++c) {   // 3 planes x 8 modes x 32 columns: 3 chunks of 16 bytes per thread
            const int chunk = tid + c * 128;
            const int p = chunk / (M3_BKM * M3_BNC / 2), rem = chunk % (M3_BKM * M3_BNC / 2);
            const int r = rem / (M3_BNC / 2), nc = (rem % (M3_BNC / 2)) * 2;
            cp_async16(b_dst + (p * M3_BKM + r) * M3_BS + nc, B3 + p * plane + (size_t)(k0 + r) * Gpad + col0 + nc, 16);
        }
    };

    double t1[4][2][2], t2[4][2][2], t3[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) t1[i][j][e] = t2[i][j][e] = t3[i][j][e] = 0.0;
#pragma unroll
    for (int s = 0; s < M3_STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }
    const int ar = lane >> 2, ak = lane & 3;              // A fragment: row = lane / 4, k = lane % 4
    const int bk = lane & 3, bn = lane >> 2;              // B fragment: k = lane % 4, n = lane / 4
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<M3_STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + M3_STAGES - 1;
            if (nk < KT) load_stage(nk % M3_STAGES, nk);
            cp_async_commit();
        }
        const double2* a_s = sA + (kt % M3_STAGES) * M3_A_STAGE + (wm * 32 + ar) * M3_AS + ak;
        const double* b_s = sB + (kt % M3_STAGES) * M3_B_STAGE + bk * M3_BS + wn * 16 + bn;
#pragma unroll
        for (int ks = 0; ks < M3_BKM / 4; ++ks) {
            double afr[4], afi[4], afs[4], bfr[2], bfi[2], bfs[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 v = a_s[i * 8 * M3_AS + ks * 4];
                afr[i] = v.x;
                afi[i] = v.y;
                afs[i] = v.x + v.y;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                bfr[j] = b_s[(0 * M3_BKM + ks * 4) * M3_BS + j * 8];
                bfi[j] = b_s[(1 * M3_BKM + ks * 4) * M3_BS + j * 8];
                bfs[j] = b_s[(2 * M3_BKM + ks * 4) * M3_BS + j * 8];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    dmma8x8x4(t1[i][j][0], t1[i][j][1], afr[i], bfr[j]);
                    dmma8x8x4(t2[i][j][0], t2[i][j][1], afi[i], bfi[j]);
                    dmma8x8x4(t3[i][j][0], t3[i][j][1], afs[i], bfs[j]);
                }
        }
    }
    cp_async_wait<0>();
    // epilogue: Re = T1 - T2, Im = T3 - T1 - T2, then (v - offset) * scale; C fragment: row = lane / 4, complex columns 2 (lane % 4) + {0, 1}
    const int cr = lane >> 2, cc = (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int col = col0 + wn * 16 + j * 8 + cc + e;
            if (col >= G) continue;
            const double2 off = *reinterpret_cast<const double2*>(offset + 2 * col);
            const double2 scl = *reinterpret_cast<const double2*>(scale + 2 * col);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t row = row0 + wm * 32 + i * 8 + cr;
                if (row >= M) continue;
                double2 v;
                v.x = ((t1[i][j][e] - t2[i][j][e]) - off.x) * scl.x;
                v.y = ((t3[i][j][e] - t1[i][j][e] - t2[i][j][e]) - off.y) * scl.y;
                C[row * G + col] = v;
            }
        }
    }
}

}  // namespace scrib200

extern "C" int scrib200_swsh_pack3m(const double* Bmat, int Kpad, int Ncpad, int n_modes, int G, double* B3, int npad, int Gpad,
                                    void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(Bmat && B3, "swsh_pack3m: null pointer");
    SCRIB200_REQUIRE(Kpad >= 2 * n_modes && Ncpad >= 2 * G, "swsh_pack3m: the packed table [%d, %d] does not hold %d modes x %d points", Kpad, Ncpad, n_modes, G);
    SCRIB200_REQUIRE(npad % M3_BKM == 0 && npad >= n_modes && Gpad % M3_BNC == 0 && Gpad >= G,
                     "swsh_pack3m: npad=%d must be a multiple of %d >= %d and Gpad=%d a multiple of %d >= %d", npad, M3_BKM, n_modes, Gpad, M3_BNC, G);
    const int64_t total = (int64_t)npad * Gpad;
    swsh_pack3m_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(Bmat, Ncpad, n_modes, G, B3, npad, Gpad);
    SCRIB200_CHECK_LAUNCH("swsh_pack3m");
    return SCRIB200_OK;
}

extern "C" int scrib200_swsh_synthesize_3m(const double* modes, int64_t n_times, int n_modes, const double* B3, int npad, int Gpad,
                                           const double* offset, const double* scale, int G, double* F, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(modes && B3 && offset && scale && F, "swsh_synthesize_3m: null pointer");
    SCRIB200_REQUIRE(n_modes > 0 && G > 0, "swsh_synthesize_3m: bad sizes n_modes=%d G=%d", n_modes, G);
    SCRIB200_REQUIRE(npad % M3_BKM == 0 && npad >= n_modes && Gpad % M3_BNC == 0 && Gpad >= G,
                     "swsh_synthesize_3m: npad=%d must be a multiple of %d >= %d and Gpad=%d a multiple of %d >= %d", npad, M3_BKM, n_modes, Gpad, M3_BNC, G);
    SCRIB200_REQUIRE(aligned16(modes) && aligned16(B3) && aligned16(F) && aligned16(offset) && aligned16(scale),
                     "swsh_synthesize_3m: pointers must be 16-byte aligned");
    if (n_times <= 0) return SCRIB200_OK;
    int band = (int)(((size_t)48 << 20) / ((size_t)3 * npad * M3_BNC * sizeof(double)));
    if (band < 1) band = 1;
    int variant = 3;                                       // ring depth: 3 stages (57.6 KB, three CTAs per SM) measured best
    if (const char* env = getenv("SCRIB200_SYNTH3M_STAGES")) variant = atoi(env);
    const int64_t max_rows = (int64_t)65535 * M3_BM;
    for (int64_t r0 = 0; r0 < n_times; r0 += max_rows) {
        const int64_t rows = n_times - r0 < max_rows ? n_times - r0 : max_rows;
        dim3 grid(Gpad / M3_BNC, (unsigned)((rows + M3_BM - 1) / M3_BM));
#define M3_LAUNCH(ST_)                                                                                                     \
    do {                                                                                                                   \
        cudaFuncSetAttribute(swsh_synth3m_kernel<ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m3_smem<ST_>());   \
        swsh_synth3m_kernel<ST_><<<grid, 128, m3_smem<ST_>(), (cudaStream_t)stream>>>(                                      \
            reinterpret_cast<const double2*>(modes) + r0 * n_modes, rows, n_modes, B3, npad, Gpad, offset, scale, G,        \
            reinterpret_cast<double2*>(F) + r0 * G, band);                                                                  \
    } while (0)
        if (variant == 2) M3_LAUNCH(2);
        else if (variant == 4) M3_LAUNCH(4);
        else M3_LAUNCH(3);
#undef M3_LAUNCH
        SCRIB200_CHECK_LAUNCH("swsh_synthesize_3m");
    }
    return SCRIB200_OK;
}
