// K3: SWSH analysis (map2salm) batched over time steps.
//
// Replaces scri/waveform_grid.py:303-307 (one spinsfast.map2salm call per time step, the first
// ell_min^2 coefficients dropped).  Huffenberger & Wandelt's algorithm (torus extension, weight
// convolution, Delta-matrix sums) is, once folded back onto [0, pi], a phi-DFT followed by a
// Clenshaw-Curtis quadrature in theta against sY_lm(theta_j, 0) (see scri_b200/_sf.py:analysis_tables):
//     f_m(theta_j) = sum_k f[j,k] E[k,m]            E = e^{-i m phi_k} / n_phi
//     a_lm         = sum_j Wt[lm, j] f_m(theta_j)   Wt = 2 pi q_j sY_lm(theta_j, 0)   (real)
//
// Fused kernel: a CTA stages T time steps of the grid ([T, n_theta*n_phi] complex, contiguous in HBM)
// and both tables in shared memory, does the DFT into shared memory, then the theta contraction, and
// writes [T, n_modes] coalesced.  The grid is read from HBM exactly once.
// Fallback for grids too large for shared memory: two passes through a workspace.
#include <stdlib.h>

#include "common.cuh"

namespace scrib200 {

__device__ __forceinline__ void lm_from_index(int idx, int ell_min, int& ell, int& m) {
    const int full = idx + ell_min * ell_min;
    int l = (int)floor(sqrt((double)full));
    while (l * l > full) --l;
    while ((l + 1) * (l + 1) <= full) ++l;
    ell = l;
    m = full - l * (l + 1);
}

__global__ void map2salm_fused_kernel(const double2* __restrict__ grid, int64_t n_times, int n_theta, int n_phi,
                                      const double2* __restrict__ E, const double* __restrict__ Wt, int ell_min,
                                      int ell_max, double2* __restrict__ out, int T) {
    extern __shared__ double2 sm2[];
    const int nm = 2 * ell_max + 1;
    const int G = n_theta * n_phi;
    const int n_modes = ell_max * (ell_max + 2) - ell_min * ell_min + 1;
    double2* sE = sm2;                         // [n_phi][nm]
    double2* sF = sE + n_phi * nm;             // [T][G]
    double2* sFm = sF + (size_t)T * G;         // [T][n_theta][nm]
    double* sW = reinterpret_cast<double*>(sFm + (size_t)T * n_theta * nm);   // [n_modes][n_theta]

    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * T;
    const int Tv = (int)((n_times - t0 < T) ? (n_times - t0) : T);

    for (int i = tid; i < n_phi * nm; i += nt) sE[i] = E[i];
    for (int i = tid; i < n_modes * n_theta; i += nt) sW[i] = Wt[i];
    const double2* src = grid + t0 * G;
    for (int i = tid; i < Tv * G; i += nt) sF[i] = src[i];
    __syncthreads();

    // phi-DFT: (tt, j, mi)
    for (int idx = tid; idx < Tv * n_theta * nm; idx += nt) {
        const int mi = idx % nm;
        const int tj = idx / nm;   // tt*n_theta + j
        const double2* f = sF + (size_t)tj * n_phi;
        double2 acc = make_double2(0.0, 0.0);
        for (int k = 0; k < n_phi; ++k) cfma(acc, f[k], sE[k * nm + mi]);
        sFm[idx] = acc;
    }
    __syncthreads();

    // theta quadrature: (tt, lm)
    for (int idx = tid; idx < Tv * n_modes; idx += nt) {
        const int tt = idx / n_modes;
        const int lm = idx - tt * n_modes;
        int ell, m;
        lm_from_index(lm, ell_min, ell, m);
        const double2* fm = sFm + (size_t)tt * n_theta * nm + (m + ell_max);
        const double* w = sW + lm * n_theta;
        double2 acc = make_double2(0.0, 0.0);
        for (int j = 0; j < n_theta; ++j) {
            const double2 v = fm[j * nm];
            acc.x = fma(w[j], v.x, acc.x);
            acc.y = fma(w[j], v.y, acc.y);
        }
        out[(t0 + tt) * n_modes + lm] = acc;
    }
}

// ---- fallback: DFT pass to a workspace [n_times, n_theta, nm], then quadrature pass
__global__ void map2salm_dft_kernel(const double2* __restrict__ grid, int64_t total /* n_times*n_theta*nm */,
                                    int n_phi, int nm, const double2* __restrict__ E, double2* __restrict__ fm) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int mi = (int)(idx % nm);
    const int64_t tj = idx / nm;
    const double2* f = grid + tj * n_phi;
    double2 acc = make_double2(0.0, 0.0);
    for (int k = 0; k < n_phi; ++k) cfma(acc, f[k], E[k * nm + mi]);
    fm[idx] = acc;
}

__global__ void map2salm_quad_kernel(const double2* __restrict__ fm, int64_t n_times, int n_theta, int ell_min,
                                     int ell_max, const double* __restrict__ Wt, double2* __restrict__ out) {
    const int nm = 2 * ell_max + 1;
    const int n_modes = ell_max * (ell_max + 2) - ell_min * ell_min + 1;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_times * n_modes) return;
    const int64_t tt = idx / n_modes;
    const int lm = (int)(idx - tt * n_modes);
    int ell, m;
    lm_from_index(lm, ell_min, ell, m);
    const double2* f = fm + tt * n_theta * nm + (m + ell_max);
    const double* w = Wt + (size_t)lm * n_theta;
    double2 acc = make_double2(0.0, 0.0);
    for (int j = 0; j < n_theta; ++j) {
        const double2 v = f[(size_t)j * nm];
        acc.x = fma(w[j], v.x, acc.x);
        acc.y = fma(w[j], v.y, acc.y);
    }
    out[idx] = acc;
}

// ---- fast path inside transform(): the grid arrives time-tiled, tile b = contiguous [G][T] block (what the
// spline remap writes).  One elected thread pulls the whole tile into shared memory with TMA bulk copies
// (cp.async.bulk + mbarrier; the second CTA resident on the SM computes meanwhile), then one thread per
// (theta ring j, time step tt) walks the ring's n_phi samples (quarter-warps read 128 contiguous bytes: no bank
// conflicts) and accumulates all m at once.  The +m / -m pair shares its four real products:
//   f (c -+ i s):  P1 = sum fr c, P2 = sum fi s, P3 = sum fi c, P4 = sum fr s;
//   f_{+m} = (P1+P2) + i(P3-P4),  f_{-m} = (P1-P2) + i(P3+P4)            - half the FMAs of a complex DFT.
// The f_m(theta_j) results overwrite the tile (after a barrier) and feed the theta quadrature.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int NP>
__global__ void __launch_bounds__(256)
map2salm_tiled_kernel(const double2* __restrict__ gridT, int64_t n_times, int n_theta, int n_phi,
                      const double2* __restrict__ trig, const double* __restrict__ Wt, int ell_min, int ell_max,
                      double2* __restrict__ out, int T, int alias) {
    extern __shared__ __align__(128) double2 sm3[];
    __shared__ __align__(8) unsigned long long mbar;
    const int L = ell_max, nm = 2 * L + 1, np1 = L + 1;
    const int G = n_theta * n_phi;
    const int n_modes = L * (L + 2) - ell_min * ell_min + 1;
    double2* sTile = sm3;                                                  // [G][T]
    double2* sFm = alias ? sTile : sTile + (size_t)G * T;                  // [T][n_theta][nm]
    const size_t fm_elems = (size_t)T * n_theta * nm, tile_elems = (size_t)G * T;
    double2* sTrig = sTile + (alias ? (fm_elems > tile_elems ? fm_elems : tile_elems) : tile_elems + fm_elems);
    double* sW = reinterpret_cast<double*>(sTrig + n_phi * np1);           // [n_modes][n_theta]
    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * T;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned total = (unsigned)(tile_elems * sizeof(double2));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(total) : "memory");
        const char* src = reinterpret_cast<const char*>(gridT + (int64_t)blockIdx.x * tile_elems);
        char* dst = reinterpret_cast<char*>(sTile);
        for (unsigned off = 0; off < total; off += 32768u) {
            const unsigned n = (total - off < 32768u) ? (total - off) : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(dst + off)),
                         "l"(src + off), "r"(n), "r"(smem_u32(&mbar))
                         : "memory");
        }
    }
    for (int i = tid; i < n_phi * np1; i += nt) sTrig[i] = trig[i];
    for (int i = tid; i < n_modes * n_theta; i += nt) sW[i] = Wt[i];
    __syncthreads();   // tables visible; mbarrier initialised before anyone polls it
    {
        unsigned ok = 0;
        while (!ok) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok)
                         : "r"(smem_u32(&mbar))
                         : "memory");
        }
    }

    const bool worker = tid < T * n_theta;
    const int j = tid / T, tt = tid - (tid / T) * T;
    for (int p0 = 0; p0 <= L; p0 += NP) {
        double acc[NP][4];
#pragma unroll
        for (int q = 0; q < NP; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0;
        if (worker) {
            const double2* src = sTile + (size_t)j * n_phi * T + tt;
#pragma unroll 5
            for (int k = 0; k < n_phi; ++k) {
                const double2 f = src[k * T];
                const double2* cs = sTrig + k * np1 + p0;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    if (p0 + q <= L) {
                        const double2 c = cs[q];
                        acc[q][0] = fma(f.x, c.x, acc[q][0]);
                        acc[q][1] = fma(f.y, c.y, acc[q][1]);
                        acc[q][2] = fma(f.y, c.x, acc[q][2]);
                        acc[q][3] = fma(f.x, c.y, acc[q][3]);
                    }
                }
            }
        }
        if (alias) __syncthreads();   // single pass: everyone is done reading the tile before it is overwritten
        if (worker) {
            double2* dst = sFm + ((size_t)tt * n_theta + j) * nm + L;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int p = p0 + q;
                if (p <= L) {
                    dst[p] = make_double2(acc[q][0] + acc[q][1], acc[q][2] - acc[q][3]);
                    if (p > 0) dst[-p] = make_double2(acc[q][0] - acc[q][1], acc[q][2] + acc[q][3]);
                }
            }
        }
    }
    __syncthreads();

    const int Tv = (int)((n_times - t0 < T) ? (n_times - t0) : T);
    for (int idx = tid; idx < Tv * n_modes; idx += nt) {
        const int t2 = idx / n_modes;
        const int lm = idx - t2 * n_modes;
        int ell, m;
        lm_from_index(lm, ell_min, ell, m);
        const double2* fm = sFm + (size_t)t2 * n_theta * nm + (m + L);
        const double* w = sW + lm * n_theta;
        double2 acc = make_double2(0.0, 0.0);
        for (int jj = 0; jj < n_theta; ++jj) {
            const double2 v = fm[jj * nm];
            acc.x = fma(w[jj], v.x, acc.x);
            acc.y = fma(w[jj], v.y, acc.y);
        }
        out[(t0 + t2) * n_modes + lm] = acc;
    }
}

// ---- DMMA variant of the tiled kernel (time tile T = 8).  For one theta ring j the phi-DFT of all 8 time steps is
// the real GEMM   P[(part, t), n] = sum_k  part(f[j, k, t]) * B[k, n],   B = (cos(m phi_k), sin(m phi_k))/n_phi, m = 1..L,
// i.e. M = 2 x 8 (Re / Im part x time), K = n_phi, N = 2 x 8 NT: FP64 tensor-core tiles m8n8k4 (SASS DMMA.8x8x4).  The
// A fragment of BOTH M tiles is one 16-byte shared-memory load straight from the [G][8] tile the spline kernel wrote
// (lane (t, kk) reads the complex sample (j, k0+kk, t): 512 contiguous-ish bytes per warp, the 4-wavefront minimum).
// The four real products of a +-m pair share the fragments (see map2salm_tiled_kernel); m = 0 is the plain ring sum,
// taken from the A fragments.  A warp owns up to two rings and keeps their accumulators in registers until every warp
// has finished reading the tile, so f_m(theta_j) can overwrite it; the theta quadrature then runs as before.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int NT>
__global__ void __launch_bounds__(640)
map2salm_dmma_kernel(const double2* __restrict__ gridT, int64_t n_times, int n_theta, int n_phi,
                     const double2* __restrict__ trig, const double* __restrict__ Wt, int ell_min, int ell_max,
                     double2* __restrict__ out) {
    constexpr int T = 8;
    constexpr int BP = 16 * NT + 8;                        // pitch of the B table (doubles): 2-way banks = the minimum
    extern __shared__ __align__(128) double2 smd[];
    __shared__ __align__(8) unsigned long long mbar;
    const int L = ell_max, nm = 2 * L + 1, np1 = L + 1;
    const int G = n_theta * n_phi;
    const int KS = (n_phi + 3) / 4;
    const int n_modes = L * (L + 2) - ell_min * ell_min + 1;
    double2* sTile = smd;                                  // [G + 4][T]  (4 zeroed pad rows: the K padding of the last ring)
    double2* sFm = sTile;                                  // [T][nm][n_theta] after the barrier
    double* sB = reinterpret_cast<double*>(sTile + (size_t)(G + 4) * T);   // [4 KS][BP]
    double* sW = sB + (size_t)4 * KS * BP;                 // [n_modes][n_theta]
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * T;
    const size_t tile_elems = (size_t)G * T;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned total = (unsigned)(tile_elems * sizeof(double2));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(total) : "memory");
        const char* src = reinterpret_cast<const char*>(gridT + (int64_t)blockIdx.x * tile_elems);
        char* dst = reinterpret_cast<char*>(sTile);
        for (unsigned off = 0; off < total; off += 32768u) {
            const unsigned n = (total - off < 32768u) ? (total - off) : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(dst + off)),
                         "l"(src + off), "r"(n), "r"(smem_u32(&mbar))
                         : "memory");
        }
    }
    for (int i = tid; i < 4 * T; i += nt) sTile[tile_elems + i] = make_double2(0.0, 0.0);
    for (int i = tid; i < 4 * KS * BP; i += nt) {
        const int k = i / BP, n = i - k * BP;
        double v = 0.0;
        if (k < n_phi && n < 16 * NT) {
            const int m = (n % (8 * NT)) + 1;
            if (m <= L) {
                const double2 cs = trig[k * np1 + m];
                v = (n < 8 * NT) ? cs.x : cs.y;
            }
        }
        sB[i] = v;
    }
    for (int i = tid; i < n_modes * n_theta; i += nt) sW[i] = Wt[i];
    __syncthreads();   // tables visible; mbarrier initialised before anyone polls it
    {
        unsigned ok = 0;
        while (!ok) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok)
                         : "r"(smem_u32(&mbar))
                         : "memory");
        }
    }

    // ---- phi-DFT on the tensor cores: rings warp, warp + nwarp
    const int tq = lane >> 2, kk = lane & 3;               // fragment coordinates: row / column t (or n), K index kk
    double acc[2][2][2 * NT][2];                           // [ring][Re/Im part][N tile][2]
    double s0[2][2];
    const double inv_nphi = trig[0].x;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        s0[r][0] = s0[r][1] = 0.0;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int n = 0; n < 2 * NT; ++n) acc[r][p][n][0] = acc[r][p][n][1] = 0.0;
        const int j = warp + r * nwarp;
        if (j < n_theta) {
            const double2* arow = sTile + ((size_t)j * n_phi + kk) * T + tq;
            const double* brow = sB + kk * BP + tq;
#pragma unroll 2
            for (int ks = 0; ks < KS; ++ks) {
                const double2 a = arow[(size_t)ks * 4 * T];
                if (4 * ks + kk < n_phi) {
                    s0[r][0] += a.x;
                    s0[r][1] += a.y;
                }
#pragma unroll
                for (int n = 0; n < 2 * NT; ++n) {
                    const double b = brow[ks * 4 * BP + n * 8];
                    dmma884(acc[r][0][n][0], acc[r][0][n][1], a.x, b);
                    dmma884(acc[r][1][n][0], acc[r][1][n][1], a.y, b);
                }
            }
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                s0[r][p] += __shfl_xor_sync(0xffffffffu, s0[r][p], 1);
                s0[r][p] += __shfl_xor_sync(0xffffffffu, s0[r][p], 2);
            }
        }
    }
    __syncthreads();   // every warp is done reading the tile: f_m(theta_j) may overwrite it
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int j = warp + r * nwarp;
        if (j < n_theta) {
            double2* frow = sFm + (size_t)tq * nm * n_theta + j;       // + mi * n_theta
            if (kk == 0) frow[(size_t)L * n_theta] = make_double2(s0[r][0] * inv_nphi, s0[r][1] * inv_nphi);
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int m = n * 8 + 2 * kk + i + 1;
                    if (m <= L) {
                        const double P1 = acc[r][0][n][i], P4 = acc[r][0][NT + n][i];
                        const double P3 = acc[r][1][n][i], P2 = acc[r][1][NT + n][i];
                        frow[(size_t)(L + m) * n_theta] = make_double2(P1 + P2, P3 - P4);
                        frow[(size_t)(L - m) * n_theta] = make_double2(P1 - P2, P3 + P4);
                    }
                }
        }
    }
    __syncthreads();

    // ---- theta quadrature: (t, lm)
    const int Tv = (int)((n_times - t0 < T) ? (n_times - t0) : T);
    for (int idx = tid; idx < Tv * n_modes; idx += nt) {
        const int t2 = idx / n_modes;
        const int lm = idx - t2 * n_modes;
        int ell, m;
        lm_from_index(lm, ell_min, ell, m);
        const double2* fm = sFm + ((size_t)t2 * nm + (m + L)) * n_theta;
        const double* w = sW + lm * n_theta;
        double2 acc2 = make_double2(0.0, 0.0);
        for (int jj = 0; jj < n_theta; ++jj) {
            const double2 v = fm[jj];
            acc2.x = fma(w[jj], v.x, acc2.x);
            acc2.y = fma(w[jj], v.y, acc2.y);
        }
        out[(t0 + t2) * n_modes + lm] = acc2;
    }
}

// ---- persistent DMMA kernel: one CTA per SM walks the time tiles with two tile buffers in shared memory, so the TMA
// bulk load of tile i+1 flies under the arithmetic of tile i, and the tables are set up once per CTA instead of once
// per 8 time steps.  Both contractions run on the FP64 tensor cores:
//   phi-DFT, folded: cos(m phi_k) = cos(m phi_{n-k}) and sin(m phi_k) = -sin(m phi_{n-k}), so with p_k = f_k + f_{n-k} and
//   q_k = f_k - f_{n-k} (k = 1 .. n/2; p_0 = f_0, q_0 = 0) the cosine sums run over p and the sine sums over q with
//   K = n/2 + 1 instead of n: 16 DMMA per ring instead of 28 at n_phi = 25.  M = t, K = k, N = m; the A fragments are
//   two 16-byte loads from the [G][8] tile the spline kernel wrote, the trig fragments live in registers for the whole
//   kernel (NT = 1) or in a conflict-free shared table (NT = 2);
//   theta quadrature, per m:  a[(l, m), (t, part)] = sum_j W[(l, m), j] f_m[t][j][part]:  M = l (rows of Wt, read through
//   L1), K = j, N = t for the Re and the Im tile - the B fragment of both is one 16-byte load of f_m(theta_j) from the
//   buffer the DFT results were written to (which is the tile buffer itself, after a barrier).
template <int NT>
__global__ void __launch_bounds__(544, 1)
map2salm_persist_kernel(const double2* __restrict__ gridT, int64_t n_times, int n_theta, int n_phi,
                        const double2* __restrict__ trig, const double* __restrict__ Wt, int ell_min, int ell_max,
                        double2* __restrict__ out) {
    constexpr int T = 8;
    constexpr int BP = 16 * NT + 4;                                          // pitch of the folded trig table: conflict free
    constexpr int KF = 4;                                                    // k-steps of the folded DFT (n_phi <= 31)
    constexpr int KQM = 8;                                                   // k-steps of the quadrature held in registers (n_theta <= 32)
    extern __shared__ __align__(128) double2 smp[];
    __shared__ __align__(8) unsigned long long mbar[2];
    const int L = ell_max, nm = 2 * L + 1, np1 = L + 1;
    const int G = n_theta * n_phi;
    const int KQ = (n_theta + 3) / 4;
    const int half = n_phi / 2;                                              // folded index k = 0 .. half
    const int n_modes = L * (L + 2) - ell_min * ell_min + 1;
    const size_t tile_elems = (size_t)G * T;
    const size_t buf_elems = (size_t)(G + 4) * T;
    // two tile buffers (TMA destinations), then f_m(theta_j) in a buffer of its own: a tile buffer is free for its refill as
    // soon as the DFT has read it, one and a half tiles of work before the refill is needed
    double2* sFm = smp + 2 * buf_elems;                                      // [T][nm][n_theta]
    double* sB = reinterpret_cast<double*>(sFm + (size_t)T * nm * n_theta);  // [4 KF][BP]: cos | sin of (m phi_k) / n_phi
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int64_t ntiles = (n_times + T - 1) / T;
    const unsigned tile_bytes = (unsigned)(tile_elems * sizeof(double2));

    auto issue = [&](int64_t tile, int b) {   // thread 0 only
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[b])), "r"(tile_bytes) : "memory");
        const char* src = reinterpret_cast<const char*>(gridT + tile * (int64_t)tile_elems);
        const unsigned dst = smem_u32(smp + (size_t)b * buf_elems);
        for (unsigned off = 0; off < tile_bytes; off += 32768u) {
            const unsigned n = (tile_bytes - off < 32768u) ? (tile_bytes - off) : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
                         "l"(src + off), "r"(n), "r"(smem_u32(&mbar[b]))
                         : "memory");
        }
    };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((int64_t)blockIdx.x < ntiles) issue(blockIdx.x, 0);
        if ((int64_t)blockIdx.x + gridDim.x < ntiles) issue((int64_t)blockIdx.x + gridDim.x, 1);
    }
    for (int i = tid; i < 4 * KF * BP; i += nt) {
        const int k = i / BP, n = i - k * BP;
        double v = 0.0;
        if (k <= half && n < 16 * NT) {
            const int m = (n % (8 * NT)) + 1;
            if (m <= L) {
                const double2 cs = trig[k * np1 + m];
                v = (n < 8 * NT) ? cs.x : cs.y;
            }
        }
        sB[i] = v;
    }
    __syncthreads();
    const int tq = lane >> 2, kk = lane & 3;
    const double inv_nphi = trig[0].x;
    const int MT = (L - ell_min + 8) / 8;                                    // 8-row blocks of l for the longest m column
    // trig fragments of this lane (B[k = kk][n = tq]) for every k-step: registers when NT = 1
    double bfrag[KF][2];
    if (NT == 1) {
#pragma unroll
        for (int ks = 0; ks < KF; ++ks) {
            bfrag[ks][0] = sB[(4 * ks + kk) * BP + tq];
            bfrag[ks][1] = sB[(4 * ks + kk) * BP + 8 + tq];
        }
    }
    // quadrature unit of this warp when every unit has a warp of its own (the usual case): its weights W[(l, m), j] are the
    // A fragments of every tile - read once, kept in registers
    const bool own_unit = (nm * MT <= nwarp);
    double wfrag[KQM];
    int u_mi = 0, u_m = 0, u_la = 0;
    bool u_valid = false, u_rowok = false;
    if (own_unit && warp < nm * MT) {
        u_mi = warp / MT;
        const int mt = warp - u_mi * MT;
        u_m = u_mi - L;
        const int am = u_m < 0 ? -u_m : u_m;
        const int lstart = (am > ell_min ? am : ell_min) + 8 * mt;
        u_valid = lstart <= L;
        u_la = lstart + tq;
        u_rowok = u_valid && u_la <= L;
#pragma unroll
        for (int kq = 0; kq < KQM; ++kq) {
            const int jj = 4 * kq + kk;
            wfrag[kq] = (u_rowok && jj < n_theta) ? Wt[(size_t)(u_la * (u_la + 1) - ell_min * ell_min + u_m) * n_theta + jj] : 0.0;
        }
    }

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        const unsigned parity = (unsigned)((it >> 1) & 1);
        {
            unsigned ok = 0;
            while (!ok) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok)
                             : "r"(smem_u32(&mbar[b])), "r"(parity)
                             : "memory");
            }
        }
        const double2* sTile = smp + (size_t)b * buf_elems;
        // ---- phi-DFT: rings warp, warp + nwarp
        double acc[2][2][2 * NT][2];                                         // [ring][Re / Im part][cos tiles | sin tiles][2]
        double s0[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            s0[r][0] = s0[r][1] = 0.0;
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int n = 0; n < 2 * NT; ++n) acc[r][p][n][0] = acc[r][p][n][1] = 0.0;
            const int j = warp + r * nwarp;
            if (j < n_theta) {
                const double2* ring = sTile + (size_t)j * n_phi * T + tq;
#pragma unroll
                for (int ks = 0; ks < KF; ++ks) {
                    const int k = 4 * ks + kk;                               // folded index and its mirror n_phi - k
                    const bool mirrored = (k >= 1 && k <= half && 2 * k != n_phi);
                    double2 a1 = make_double2(0.0, 0.0), a2 = make_double2(0.0, 0.0);
                    if (k <= half) a1 = ring[k * T];
                    if (mirrored) a2 = ring[(n_phi - k) * T];
                    const double2 pp = make_double2(a1.x + a2.x, a1.y + a2.y);
                    const double2 qq = mirrored ? make_double2(a1.x - a2.x, a1.y - a2.y) : make_double2(0.0, 0.0);
                    s0[r][0] += pp.x;
                    s0[r][1] += pp.y;
#pragma unroll
                    for (int n = 0; n < NT; ++n) {
                        const double bc = (NT == 1) ? bfrag[ks][0] : sB[(4 * ks + kk) * BP + n * 8 + tq];
                        const double bs = (NT == 1) ? bfrag[ks][1] : sB[(4 * ks + kk) * BP + (NT + n) * 8 + tq];
                        dmma884(acc[r][0][n][0], acc[r][0][n][1], pp.x, bc);
                        dmma884(acc[r][1][n][0], acc[r][1][n][1], pp.y, bc);
                        dmma884(acc[r][0][NT + n][0], acc[r][0][NT + n][1], qq.x, bs);
                        dmma884(acc[r][1][NT + n][0], acc[r][1][NT + n][1], qq.y, bs);
                    }
                }
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    s0[r][p] += __shfl_xor_sync(0xffffffffu, s0[r][p], 1);
                    s0[r][p] += __shfl_xor_sync(0xffffffffu, s0[r][p], 2);
                }
            }
        }
        __syncthreads();   // every warp is done reading the tile (and, from the previous tile, f_m): refill and overwrite
        if (tid == 0) {
            const int64_t nxt = tile + 2 * (int64_t)gridDim.x;
            if (nxt < ntiles) issue(nxt, b);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int j = warp + r * nwarp;
            if (j < n_theta) {
                double2* frow = sFm + (size_t)tq * nm * n_theta + j;       // + mi * n_theta
                if (kk == 0) frow[(size_t)L * n_theta] = make_double2(s0[r][0] * inv_nphi, s0[r][1] * inv_nphi);
#pragma unroll
                for (int n = 0; n < NT; ++n)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int m = n * 8 + 2 * kk + i + 1;
                        if (m <= L) {
                            const double P1 = acc[r][0][n][i], P4 = acc[r][0][NT + n][i];
                            const double P3 = acc[r][1][n][i], P2 = acc[r][1][NT + n][i];
                            frow[(size_t)(L + m) * n_theta] = make_double2(P1 + P2, P3 - P4);
                            frow[(size_t)(L - m) * n_theta] = make_double2(P1 - P2, P3 + P4);
                        }
                    }
            }
        }
        __syncthreads();
        // ---- theta quadrature on the tensor cores: units (mi, 8-row block of l)
        const int64_t t0 = tile * T;
        if (own_unit) {
            if (u_valid) {
                const double2* fcol = sFm + ((size_t)tq * nm + u_mi) * n_theta + kk;   // B fragment: column t = tq, rows j
                double cre[2] = {0.0, 0.0}, cim[2] = {0.0, 0.0};
#pragma unroll
                for (int kq = 0; kq < KQM; ++kq) {
                    if (kq < KQ) {
                        const double2 bv = (4 * kq + kk < n_theta) ? fcol[4 * kq] : make_double2(0.0, 0.0);
                        dmma884(cre[0], cre[1], wfrag[kq], bv.x);
                        dmma884(cim[0], cim[1], wfrag[kq], bv.y);
                    }
                }
                if (u_rowok) {
                    const int lm = u_la * (u_la + 1) - ell_min * ell_min + u_m;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int64_t tt = t0 + 2 * kk + i;
                        if (tt < n_times) out[tt * n_modes + lm] = make_double2(cre[i], cim[i]);
                    }
                }
            }
        } else {
            for (int unit = warp; unit < nm * MT; unit += nwarp) {
                const int mi = unit / MT, mt = unit - mi * MT;
                const int m = mi - L;
                const int am = m < 0 ? -m : m;
                const int lstart = (am > ell_min ? am : ell_min) + 8 * mt;
                if (lstart > L) continue;
                const int la = lstart + tq;                                  // my row of the A fragment
                const bool rowok = la <= L;
                const double* wrow = Wt + (size_t)(la * (la + 1) - ell_min * ell_min + m) * n_theta;
                const double2* fcol = sFm + ((size_t)tq * nm + mi) * n_theta;
                double cre[2] = {0.0, 0.0}, cim[2] = {0.0, 0.0};
#pragma unroll 2
                for (int kq = 0; kq < KQ; ++kq) {
                    const int jj = 4 * kq + kk;
                    const bool jok = jj < n_theta;
                    const double av = (rowok && jok) ? __ldg(wrow + jj) : 0.0;
                    const double2 bv = jok ? fcol[jj] : make_double2(0.0, 0.0);
                    dmma884(cre[0], cre[1], av, bv.x);
                    dmma884(cim[0], cim[1], av, bv.y);
                }
                if (rowok) {
                    const int lm = la * (la + 1) - ell_min * ell_min + m;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int64_t tt = t0 + 2 * kk + i;
                        if (tt < n_times) out[tt * n_modes + lm] = make_double2(cre[i], cim[i]);
                    }
                }
            }
        }
        // (no barrier here: the next tile's DFT touches only its own tile buffer; f_m is overwritten after the next barrier)
    }
}

static size_t persist_smem(int n_theta, int n_phi, int ell_max, int NT) {
    return (2 * ((size_t)n_theta * n_phi + 4) * 8 + (size_t)8 * (2 * ell_max + 1) * n_theta) * sizeof(double2) +
           4 * 4 * (16 * NT + 4) * sizeof(double);
}

static size_t dmma_smem(int n_theta, int n_phi, int ell_min, int ell_max, int NT) {
    const size_t n_modes = (size_t)ell_max * (ell_max + 2) - (size_t)ell_min * ell_min + 1;
    const size_t KS = (n_phi + 3) / 4;
    return ((size_t)n_theta * n_phi + 4) * 8 * sizeof(double2) + (4 * KS * (16 * NT + 8) + n_modes * n_theta) * sizeof(double);
}

constexpr int TILED_NP = 9;

static size_t gmajor_smem(int T, int n_theta, int n_phi, int ell_min, int ell_max) {
    const size_t nm = 2 * ell_max + 1;
    const size_t n_modes = (size_t)ell_max * (ell_max + 2) - (size_t)ell_min * ell_min + 1;
    const size_t tile = (size_t)n_theta * n_phi * T, fm = (size_t)T * n_theta * nm;
    const bool alias = (ell_max + 1 <= TILED_NP);
    const size_t data = alias ? (tile > fm ? tile : fm) : tile + fm;
    return (data + (size_t)n_phi * (ell_max + 1)) * sizeof(double2) + n_modes * n_theta * sizeof(double);
}

static size_t fused_smem(int T, int n_theta, int n_phi, int ell_min, int ell_max) {
    const size_t nm = 2 * ell_max + 1;
    const size_t n_modes = (size_t)ell_max * (ell_max + 2) - (size_t)ell_min * ell_min + 1;
    return (n_phi * nm + (size_t)T * n_theta * n_phi + (size_t)T * n_theta * nm) * sizeof(double2) +
           n_modes * n_theta * sizeof(double);
}

}  // namespace scrib200

extern "C" size_t scrib200_map2salm_workspace_bytes(int64_t n_times, int n_theta, int n_phi, int ell_max) {
    using namespace scrib200;
    // only the fallback path needs it; ell_min = 0 gives the larger table
    if (fused_smem(1, n_theta, n_phi, 0, ell_max) <= 200 * 1024) return 0;
    return (size_t)n_times * n_theta * (2 * ell_max + 1) * sizeof(double2);
}

extern "C" int scrib200_map2salm(const double* grid, int64_t n_times, int n_theta, int n_phi, const double* E,
                                 const double* Wt, int ell_min, int ell_max, double* out, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(grid && E && Wt && out, "map2salm: null pointer");
    SCRIB200_REQUIRE(n_theta >= 2 && n_phi >= 1, "map2salm: bad grid %d x %d", n_theta, n_phi);
    SCRIB200_REQUIRE(ell_min >= 0 && ell_max >= ell_min, "map2salm: bad ell range [%d, %d]", ell_min, ell_max);
    SCRIB200_REQUIRE(aligned16(grid) && aligned16(E) && aligned16(out), "map2salm: pointers must be 16-byte aligned");
    if (n_times <= 0) return SCRIB200_OK;
    const int nm = 2 * ell_max + 1;
    const int n_modes = ell_max * (ell_max + 2) - ell_min * ell_min + 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (fused_smem(1, n_theta, n_phi, ell_min, ell_max) <= 200 * 1024) {
        int T = 1;
        while (T < 8 && fused_smem(T + 1, n_theta, n_phi, ell_min, ell_max) <= 100 * 1024) ++T;
        const size_t smem = fused_smem(T, n_theta, n_phi, ell_min, ell_max);
        cudaFuncSetAttribute(map2salm_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int64_t blocks = (n_times + T - 1) / T;
        map2salm_fused_kernel<<<(unsigned)blocks, 256, smem, st>>>(
            reinterpret_cast<const double2*>(grid), n_times, n_theta, n_phi, reinterpret_cast<const double2*>(E), Wt,
            ell_min, ell_max, reinterpret_cast<double2*>(out), T);
        SCRIB200_CHECK_LAUNCH("map2salm(fused)");
        return SCRIB200_OK;
    }
    const size_t need = (size_t)n_times * n_theta * nm * sizeof(double2);
    SCRIB200_REQUIRE(workspace && workspace_bytes >= need, "map2salm: workspace too small (%zu < %zu)", workspace_bytes,
                     need);
    double2* fm = reinterpret_cast<double2*>(workspace);
    const int64_t total = n_times * n_theta * nm;
    map2salm_dft_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const double2*>(grid), total,
                                                                        n_phi, nm, reinterpret_cast<const double2*>(E),
                                                                        fm);
    SCRIB200_CHECK_LAUNCH("map2salm(dft)");
    const int64_t total2 = n_times * n_modes;
    map2salm_quad_kernel<<<(unsigned)((total2 + 255) / 256), 256, 0, st>>>(fm, n_times, n_theta, ell_min, ell_max, Wt,
                                                                          reinterpret_cast<double2*>(out));
    SCRIB200_CHECK_LAUNCH("map2salm(quad)");
    return SCRIB200_OK;
}

extern "C" int scrib200_map2salm_tile_size(int n_theta, int n_phi, int ell_min, int ell_max) {
    using namespace scrib200;
    // largest time tile T (power of two <= 8) whose tables fit a CTA at >= 2 CTAs per SM; 0 = use scrib200_map2salm
    for (int T = 8; T >= 2; T >>= 1)
        if (T * n_theta <= 256 && gmajor_smem(T, n_theta, n_phi, ell_min, ell_max) <= 110 * 1024) return T;
    return 0;
}

extern "C" int scrib200_map2salm_tiled(const double* gridT, int tile, int64_t n_times, int n_theta, int n_phi,
                                       const double* trig, const double* Wt, int ell_min, int ell_max, double* out,
                                       void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(gridT && trig && Wt && out, "map2salm_tiled: null pointer");
    SCRIB200_REQUIRE(n_theta >= 2 && n_phi >= 1, "map2salm_tiled: bad grid %d x %d", n_theta, n_phi);
    SCRIB200_REQUIRE(ell_min >= 0 && ell_max >= ell_min, "map2salm_tiled: bad ell range [%d, %d]", ell_min, ell_max);
    SCRIB200_REQUIRE(aligned16(gridT) && aligned16(trig) && aligned16(out), "map2salm_tiled: pointers must be 16-byte aligned");
    if (n_times <= 0) return SCRIB200_OK;
    const int T = tile;
    {
        // tensor-core path: time tile 8, up to two rings per warp, m = 1..ell_max in NT blocks of 8
        const int NT = (ell_max + 7) / 8;
        const int nwarp = (n_theta + 1) / 2;
        const size_t smem_d = dmma_smem(n_theta, n_phi, ell_min, ell_max, NT);
        const bool disabled = getenv("SCRIB200_ANALYSIS_SCALAR") != nullptr;
        const size_t smem_p = persist_smem(n_theta, n_phi, ell_max, NT);
        if (!disabled && !getenv("SCRIB200_ANALYSIS_NONPERSISTENT") && T == 8 && ell_max >= 1 && NT <= 2 && nwarp <= 20 &&
            2 * ell_max + 1 <= n_phi && 2 * ell_max + 1 <= n_theta + 0 * n_phi && n_phi <= 31 && n_theta <= 32 && nwarp <= 17 && smem_p <= 225 * 1024) {
            static int n_sm = 0;
            if (n_sm == 0) {
                int dev = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
            }
            const int64_t ntiles = (n_times + T - 1) / T;
            const unsigned blocks = (unsigned)(ntiles < n_sm ? ntiles : n_sm);
            // warps: at least one per two rings (the DFT keeps two rings' accumulators) and, if it fits, one per quadrature
            // unit (2 ell_max + 1 columns of m), so that neither phase needs a second, half-empty round
            int pw = nwarp < 4 ? 4 : nwarp;
            const int units = (2 * ell_max + 1) * ((ell_max - ell_min + 8) / 8);
            if (units > pw) pw = units < 17 ? units : 17;
            if (const char* env = getenv("SCRIB200_ANALYSIS_WARPS")) pw = atoi(env) >= nwarp && atoi(env) <= 17 ? atoi(env) : pw;
            const int threads = 32 * pw;
            if (NT == 1) {
                cudaFuncSetAttribute(map2salm_persist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p);
                map2salm_persist_kernel<1><<<blocks, threads, smem_p, (cudaStream_t)stream>>>(
                    reinterpret_cast<const double2*>(gridT), n_times, n_theta, n_phi, reinterpret_cast<const double2*>(trig), Wt,
                    ell_min, ell_max, reinterpret_cast<double2*>(out));
            } else {
                cudaFuncSetAttribute(map2salm_persist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p);
                map2salm_persist_kernel<2><<<blocks, threads, smem_p, (cudaStream_t)stream>>>(
                    reinterpret_cast<const double2*>(gridT), n_times, n_theta, n_phi, reinterpret_cast<const double2*>(trig), Wt,
                    ell_min, ell_max, reinterpret_cast<double2*>(out));
            }
            SCRIB200_CHECK_LAUNCH("map2salm_tiled(persistent dmma)");
            return SCRIB200_OK;
        }
        if (!disabled && T == 8 && ell_max >= 1 && NT <= 2 && nwarp <= 20 && 2 * ell_max + 1 <= n_phi && smem_d <= 200 * 1024) {
            const int64_t blocks = (n_times + T - 1) / T;
            const int threads = 32 * (nwarp < 4 ? 4 : nwarp);
            if (NT == 1) {
                cudaFuncSetAttribute(map2salm_dmma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d);
                map2salm_dmma_kernel<1><<<(unsigned)blocks, threads, smem_d, (cudaStream_t)stream>>>(
                    reinterpret_cast<const double2*>(gridT), n_times, n_theta, n_phi, reinterpret_cast<const double2*>(trig), Wt,
                    ell_min, ell_max, reinterpret_cast<double2*>(out));
            } else {
                cudaFuncSetAttribute(map2salm_dmma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d);
                map2salm_dmma_kernel<2><<<(unsigned)blocks, threads, smem_d, (cudaStream_t)stream>>>(
                    reinterpret_cast<const double2*>(gridT), n_times, n_theta, n_phi, reinterpret_cast<const double2*>(trig), Wt,
                    ell_min, ell_max, reinterpret_cast<double2*>(out));
            }
            SCRIB200_CHECK_LAUNCH("map2salm_tiled(dmma)");
            return SCRIB200_OK;
        }
    }
    const size_t smem = gmajor_smem(T, n_theta, n_phi, ell_min, ell_max);
    SCRIB200_REQUIRE(T >= 2 && T * n_theta <= 256 && smem <= 200 * 1024,
                     "map2salm_tiled: tile %d with grid %d x %d, ell_max=%d does not fit one CTA (use scrib200_map2salm)", T,
                     n_theta, n_phi, ell_max);
    int threads = ((T * n_theta + 31) / 32) * 32;
    if (threads < 128) threads = 128;
    cudaFuncSetAttribute(map2salm_tiled_kernel<TILED_NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t blocks = (n_times + T - 1) / T;
    map2salm_tiled_kernel<TILED_NP><<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2*>(gridT), n_times, n_theta, n_phi, reinterpret_cast<const double2*>(trig), Wt,
        ell_min, ell_max, reinterpret_cast<double2*>(out), T, (ell_max + 1 <= TILED_NP) ? 1 : 0);
    SCRIB200_CHECK_LAUNCH("map2salm_tiled");
    return SCRIB200_OK;
}
