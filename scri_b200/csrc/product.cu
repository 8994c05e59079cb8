// K9: modes of the pointwise product of two spin-weighted fields, separable and fused - the (theta, phi) grids never
// exist in HBM.
//
// Replaces the chain of scri/modes_time_series.py:177-193 (ModesTimeSeries.grid_multiply):
//   spinsfast.salm2map(a1, s1) ; spinsfast.salm2map(a2, s2) ; product ; spinsfast.map2salm(., s1+s2, L_w)[: (L_out+1)^2]
// on spinsfast's regular grid theta_j = pi j/(n_theta-1), phi_k = 2 pi k/n_phi, where sY_lm(theta_j, phi_k) =
// lambda_lm(theta_j) e^{i m phi_k} separates.  Per time step and ring j:
//   (A) F1_m(j) = sum_l lambda1_lm(j) a1_lm,  F2_m(j) likewise      - Wigner-d contraction over l, FP64 DMMA,
//                                                                     rows = 8 rings, columns = 4 time steps x (re, im)
//   (B) P_M(j)  = sum_{m1+m2 = M (mod n_phi)} F1_m1(j) F2_m2(j)      - what phi-synthesis, product and the phi-DFT of the
//                                                                     analysis amount to (DFT convolution theorem, the
//                                                                     aliases of an undersampled n_phi included); FP64 FMA
//   (C) out_lM += W_lM(j) P_M(j)                                     - theta quadrature (Clenshaw-Curtis weights folded
//                                                                     into W), FP64 DMMA, accumulators in registers
// One persistent CTA per SM owns 4 consecutive time steps: their modes are staged once in shared memory in m-major order
// and the CTA walks the rings in chunks of 8; the tables (fragment-ordered by the host, L2 resident) stream through
// registers.  Algorithmic HBM traffic: 16 (n1 + n2 + n_out) bytes per time step.
#include <cooperative_groups.h>

#include <type_traits>

#include "common.cuh"

namespace scrib200 {

struct ProductParams {
    const double2* a1;
    const double2* a2;
    double2* out;
    int64_t n_times;
    int n1, n2, n_out;
    const int* perm1;       // [n1]  double offset in shared memory of mode idx of field 1 (m-major, padded to 4 per m)
    const int* perm2;       // [n2]
    const unsigned* ctl;    // [n_ctl] stage A: first entry of every warp's stream (nwarps + 1 entries), then the streams: per k-step
    int n_steps;            //   (global k-step index | F entry << 16 | last-of-task << 31), padded per warp to the ring depth; n_steps = n_ctl
    const double* lamfrag;  // [n_chunks, lam_stride] A fragments of stage A
    int64_t lam_stride;
    const int2* tiles;      // [n_tiles] (M + L_out, l0)
    int n_tiles;
    const double* wtfrag;   // [n_chunks, n_tiles, 2, 32] A fragments of stage C
    int64_t wt_stride;
    int ell1, ell2, L_out, n_phi, n_chunks, qmax;
    int szA;                // doubles of the two staged mode tiles (also the output staging area)
    int offF1, offF2;       // double offsets of the F1 / P buffer and the F2 buffer
    int stage_out;
    int smem_doubles;       // data region; the control words of stage A follow it
    int skip;               // development: bit 0 / 1 / 2 skips stage A / B / C (timing only; 0 in production)
};

__device__ __forceinline__ void dmma_p(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16_p(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}

constexpr int PRODUCT_T = 4;   // time steps per CTA pass: 4 x (re, im) = the 8 columns of one DMMA tile

// GM = consecutive M per convolution warp, MAXT = output tiles per warp, DA = table fragments in flight per lane in stage A.
// Two shapes are instantiated: 16 warps x 5 M (ell_out <= 39) and 8 warps x 9 M; the host picks by table size.
template <int GM, int MAXT, int DA, int MAXTHREADS>
__global__ void __launch_bounds__(MAXTHREADS, 1)
modes_product_kernel(const ProductParams p) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
    double2* sF1 = reinterpret_cast<double2*>(sm + p.offF1);
    const double2* sF2 = reinterpret_cast<const double2*>(sm + p.offF2);
    const int n_mout = 2 * p.L_out + 1;
    const int64_t n_tg = (p.n_times + PRODUCT_T - 1) / PRODUCT_T;
    const int fr = (lane & 3) * 8 + (lane >> 2);   // B fragment of an [k = 4 rows][8 columns] block stored 8 doubles per row

    for (int i = p.offF1 / 2 + tid; i < p.smem_doubles / 2; i += nthreads) reinterpret_cast<double2*>(sm)[i] = make_double2(0.0, 0.0);   // guards of F2
    unsigned* sctl = reinterpret_cast<unsigned*>(sm + p.smem_doubles);
    for (int i = tid; i < p.n_steps; i += nthreads) sctl[i] = __ldg(p.ctl + i);
    // this warp's output tiles (the same for every chunk and time group; the table is padded with empty tiles)
    int tile_m[MAXT];
#pragma unroll
    for (int s = 0; s < MAXT; ++s) tile_m[s] = __ldg(p.tiles + warp + s * nwarps).x * 64;

    for (int64_t tg = blockIdx.x; tg < n_tg; tg += gridDim.x) {
        const int64_t t0 = tg * PRODUCT_T;
        const int nt = (int)min((int64_t)PRODUCT_T, p.n_times - t0);
        // ---- stage the modes of these time steps, m-major: sm[perm[idx] + 2 t + (re, im)]
        for (int i = tid; i < p.szA / 2; i += nthreads) reinterpret_cast<double2*>(sm)[i] = make_double2(0.0, 0.0);
        __syncthreads();
        for (int e = tid; e < nt * p.n1; e += nthreads) {   // 16-byte asynchronous copies: all in flight at once
            int t = e / p.n1, idx = e - t * p.n1;
            cp_async16_p(sm + __ldg(p.perm1 + idx) + 2 * t, p.a1 + t0 * p.n1 + e);
        }
        for (int e = tid; e < nt * p.n2; e += nthreads) {
            int t = e / p.n2, idx = e - t * p.n2;
            cp_async16_p(sm + __ldg(p.perm2 + idx) + 2 * t, p.a2 + t0 * p.n2 + e);
        }
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();

        double acc[MAXT][2];
#pragma unroll
        for (int s = 0; s < MAXT; ++s) acc[s][0] = acc[s][1] = 0.0;

        for (int c = 0; c < p.n_chunks; ++c) {
            // ---- (A) theta synthesis of both fields for rings 8c .. 8c+7.  Every warp walks its own stream of k-steps
            //      (its share of the (field, m) tasks back to back); a ring of DA table fragments per lane is in flight so
            //      the L2 latency hides behind ~DA DMMA slots whatever the length of the individual tasks
            if (!(p.skip & 1)) {
                const double* lf = p.lamfrag + (int64_t)c * p.lam_stride + lane;
                const int s0 = (int)sctl[warp], n = (int)sctl[warp + 1] - s0;   // n is a multiple of DA (padded with no-op steps)
                const unsigned* ctl = sctl + s0;
                double fq[DA];
                unsigned cq[DA];
#pragma unroll
                for (int d = 0; d < DA; ++d) {
                    cq[d] = ctl[d];
                    fq[d] = __ldg(lf + (cq[d] & 0xffffu) * 32);
                }
                double c0[3] = {0.0, 0.0, 0.0}, c1[3] = {0.0, 0.0, 0.0};   // three tasks side by side: stream position i feeds chain i % 3
                for (int base = 0; base < n; base += DA) {
                    ctl += DA;
#pragma unroll
                    for (int d = 0; d < DA; ++d) {   // branch-free: the scheduler batches the shared-memory operand loads
                        const unsigned u = cq[d];
                        dmma_p(c0[d % 3], c1[d % 3], fq[d], sm[(u & 0xffffu) * 32 + fr]);
                        if (u >> 31) {   // F_m[item = ring*4 + t] of this chain's (field, m) is complete
                            *reinterpret_cast<double2*>(sm + p.offF1 + ((u >> 16) & 0x7fffu) * 64 + 2 * lane) = make_double2(c0[d % 3], c1[d % 3]);
                            c0[d % 3] = c1[d % 3] = 0.0;
                        }
                        cq[d] = ctl[d];   // next turn of the ring (past the end of the stream: harmless in-range loads)
                        fq[d] = __ldg(lf + (cq[d] & 0xffffu) * 32);
                    }
                }
            }
            __syncthreads();

            // ---- (B) convolution over m for the 32 (ring, time) items of the chunk: lane = item, warp = GM consecutive M
            double2 pacc[GM];
#pragma unroll
            for (int q = 0; q < GM; ++q) pacc[q] = make_double2(0.0, 0.0);
            const int gq = (int)sctl[nwarps + 1 + warp];   // my group of GM consecutive M (balanced over the SM sub-partitions by the host), -1: none
            if (gq >= 0 && !(p.skip & 2)) {
                const int M0 = -p.L_out + GM * gq;
                const double2* f1 = sF1 + lane;
                const double2* f2 = sF2 + lane;
                const int l2 = p.ell2;
                for (int q = -p.qmax; q <= p.qmax; ++q) {
                    const int Me = M0 + q * p.n_phi;
                    const int lo = max(-p.ell1, Me - l2), hi = min(p.ell1, Me + GM - 1 + l2);
                    if (lo > hi) continue;
                    // F2 carries zero guard entries (GM + 2 below, GM - 1 above), so its loads need no range checks
                    double2 y[GM + 3];   // y[k] <-> m2 = Me - m1b - 3 + k
                    const double2* f2q = f2 + (Me + l2) * 32;
#pragma unroll
                    for (int k = 0; k < GM - 1; ++k) y[k] = f2q[(1 + k - lo) * 32];
                    auto block = [&](int m1b, auto tail) {
#pragma unroll
                        for (int k = GM - 2; k >= 0; --k) y[k + 4] = y[k];
#pragma unroll
                        for (int k = 0; k < 4; ++k) y[k] = f2q[(k - 3 - m1b) * 32];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            double2 x = make_double2(0.0, 0.0);
                            if (!decltype(tail)::value || m1b + i <= hi) x = f1[(m1b + i + p.ell1) * 32];
#pragma unroll
                            for (int qq = 0; qq < GM; ++qq) cfma(pacc[qq], x, y[qq - i + 3]);
                        }
                    };
                    int m1b = lo;
#pragma unroll 2
                    for (; m1b + 3 <= hi; m1b += 4) block(m1b, std::false_type{});
                    if (m1b <= hi) block(m1b, std::true_type{});
                }
            }
            // quadrature fragments of this chunk: issued now, consumed after the two barriers below
            double2 w[MAXT];
            {
                const double2* wf = reinterpret_cast<const double2*>(p.wtfrag + (int64_t)c * p.wt_stride) + lane;
#pragma unroll
                for (int s = 0; s < MAXT; ++s) w[s] = __ldg(wf + (warp + s * nwarps) * 32);
            }
            __syncthreads();
            if (gq >= 0) {
#pragma unroll
                for (int qq = 0; qq < GM; ++qq)
                    if (GM * gq + qq < n_mout) sF1[(GM * gq + qq) * 32 + lane] = pacc[qq];
            }
            __syncthreads();

            // ---- (C) theta quadrature: out[l, M] += sum over the 8 rings of W[lM, ring] P_M[ring]
            if (!(p.skip & 4))
#pragma unroll
            for (int s = 0; s < MAXT; ++s) {
                const double* bsrc = sm + p.offF1 + tile_m[s] + fr;
                dmma_p(acc[s][0], acc[s][1], w[s].x, bsrc[0]);
                dmma_p(acc[s][0], acc[s][1], w[s].y, bsrc[32]);
            }
            __syncthreads();
        }

        // ---- modes of the product: C fragment row = l0 + lane/4, columns (t = lane%4, re/im)
        {
            const int t = lane & 3;
            double2* so = reinterpret_cast<double2*>(sm);
#pragma unroll
            for (int s = 0; s < MAXT; ++s) {
                const int2 tl = __ldg(p.tiles + warp + s * nwarps);
                const int l = tl.y + (lane >> 2), M = tl.x - p.L_out;
                if (l <= p.L_out && t < nt) {
                    const int idx = l * (l + 1) + M;
                    if (p.stage_out) so[t * p.n_out + idx] = make_double2(acc[s][0], acc[s][1]);
                    else p.out[(t0 + t) * p.n_out + idx] = make_double2(acc[s][0], acc[s][1]);
                }
            }
            if (p.stage_out) {
                __syncthreads();
                for (int e = tid; e < nt * p.n_out; e += nthreads) p.out[t0 * p.n_out + e] = so[e];
            }
            __syncthreads();
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Separable synthesis on spinsfast's regular grid, theta stage: F_m(theta_j) = sum_l lambda_lm(theta_j) a_lm for every ring
// j and every m - stage (A) of the product kernel for one field, written to HBM as [time][ring][m] complex.  The phi stage
// f(theta_j, phi_k) = sum_m F_m(theta_j) e^{i m phi_k} is then an ordinary real GEMM over rows (time, ring)
// (scrib200_swsh_synthesize with the DFT table as its operand).  Replaces spinsfast.salm2map (scri/modes_time_series.py:177-182,
// scri/asymptotic_bondi_data/transformations.py) at 8 n_theta ((l_max+1)^2 / 2 + n_phi (2 l_max + 1)) flops per step
// instead of the dense 8 n_theta n_phi (l_max+1)^2.
struct ThetaParams {
    const double2* a;
    int n;
    const int* perm;
    const unsigned* ctl;
    int n_ctl;
    const double* lamfrag;
    int64_t lam_stride;
    double2* out;           // [n_times, n_theta, nm]
    int64_t n_times;
    int n_theta, nm, n_chunks, szA, smem_doubles;
};

constexpr int THETA_FSTRIDE = 66;   // doubles per F_m entry (32 items x (re, im) + 2): the transposed read below is conflict free

template <int DA>
__global__ void __launch_bounds__(256, 2)
theta_synth_kernel(const ThetaParams p) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = 8;
    const int fr = (lane & 3) * 8 + (lane >> 2);
    const int64_t n_tg = (p.n_times + PRODUCT_T - 1) / PRODUCT_T;
    double* sF = sm + p.szA;
    unsigned* sctl = reinterpret_cast<unsigned*>(sm + p.smem_doubles);
    for (int i = tid; i < p.n_ctl; i += 256) sctl[i] = __ldg(p.ctl + i);
    for (int i = tid; i < p.smem_doubles / 2; i += 256) reinterpret_cast<double2*>(sm)[i] = make_double2(0.0, 0.0);
    __syncthreads();
    const int s0 = (int)sctl[warp], ns = (int)sctl[warp + 1] - s0;

    for (int64_t tg = blockIdx.x; tg < n_tg; tg += gridDim.x) {
        const int64_t t0 = tg * PRODUCT_T;
        const int nt = (int)min((int64_t)PRODUCT_T, p.n_times - t0);
        if (nt < PRODUCT_T) {
            for (int i = tid; i < p.szA / 2; i += 256) reinterpret_cast<double2*>(sm)[i] = make_double2(0.0, 0.0);
            __syncthreads();
        }
        for (int e = tid; e < nt * p.n; e += 256) {
            int t = e / p.n, idx = e - t * p.n;
            cp_async16_p(sm + __ldg(p.perm + idx) + 2 * t, p.a + t0 * p.n + e);
        }
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        for (int c = 0; c < p.n_chunks; ++c) {
            {
                const double* lf = p.lamfrag + (int64_t)c * p.lam_stride + lane;
                const unsigned* ctl = sctl + s0;
                double fq[DA];
                unsigned cq[DA];
#pragma unroll
                for (int d = 0; d < DA; ++d) {
                    cq[d] = ctl[d];
                    fq[d] = __ldg(lf + (cq[d] & 0xffffu) * 32);
                }
                double c0[3] = {0.0, 0.0, 0.0}, c1[3] = {0.0, 0.0, 0.0};
                for (int base = 0; base < ns; base += DA) {
                    ctl += DA;
#pragma unroll
                    for (int d = 0; d < DA; ++d) {
                        const unsigned u = cq[d];
                        dmma_p(c0[d % 3], c1[d % 3], fq[d], sm[(u & 0xffffu) * 32 + fr]);
                        if (u >> 31) {
                            *reinterpret_cast<double2*>(sF + ((u >> 16) & 0x7fffu) * THETA_FSTRIDE + 2 * lane) = make_double2(c0[d % 3], c1[d % 3]);
                            c0[d % 3] = c1[d % 3] = 0.0;
                        }
                        cq[d] = ctl[d];
                        fq[d] = __ldg(lf + (cq[d] & 0xffffu) * 32);
                    }
                }
            }
            __syncthreads();
            // transposed, coalesced store: item = (ring r, step t) = r * 4 + t; lanes run over m
            for (int item = warp; item < 32; item += NW) {
                const int r = item >> 2, t = item & 3, ring = 8 * c + r;
                if (ring < p.n_theta && t < nt) {
                    double2* dst = p.out + ((t0 + t) * p.n_theta + ring) * (int64_t)p.nm;
                    for (int m = lane; m < p.nm; m += 32) dst[m] = *reinterpret_cast<const double2*>(sF + m * THETA_FSTRIDE + 2 * item);
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Separable analysis on the regular grid, theta stage: a_lM = sum_j W_lM(theta_j) P_M(theta_j) - stage (C) of the product
// kernel with P read from HBM, where P[t, j, M] = (1/n_phi) sum_k f(theta_j, phi_k) e^{-i M phi_k} comes from the phi-DFT GEMM
// over the rows (time, ring).  Together they replace spinsfast.map2salm (scri/waveform_grid.py:303-307,
// scri/modes_time_series.py:188) for grids too large for the shared-memory tile kernels.
struct QuadParams {
    const double2* P;       // [n_times, n_theta, nm]
    double2* out;           // [n_times, n_out]
    int64_t n_times;
    int n_theta, nm, n_chunks, n_out, ell_min2, L;
    const int2* tiles;      // [8 * MAXT] (M + L, l0), padded with empty tiles
    const double* wtfrag;   // [n_chunks, 8 * MAXT * 64]
    int64_t wt_stride;
};

template <int MAXT>
__global__ void __launch_bounds__(256, 2)
theta_quad_kernel(const QuadParams p) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = (lane & 3) * 8 + (lane >> 2);
    const int64_t n_tg = (p.n_times + PRODUCT_T - 1) / PRODUCT_T;
    for (int i = tid; i < p.nm * THETA_FSTRIDE / 2; i += 256) reinterpret_cast<double2*>(sm)[i] = make_double2(0.0, 0.0);
    int tile_m[MAXT];
#pragma unroll
    for (int s = 0; s < MAXT; ++s) tile_m[s] = __ldg(p.tiles + warp + s * 8).x * THETA_FSTRIDE;
    __syncthreads();
    for (int64_t tg = blockIdx.x; tg < n_tg; tg += gridDim.x) {
        const int64_t t0 = tg * PRODUCT_T;
        const int nt = (int)min((int64_t)PRODUCT_T, p.n_times - t0);
        double acc[MAXT][2];
#pragma unroll
        for (int s = 0; s < MAXT; ++s) acc[s][0] = acc[s][1] = 0.0;
        for (int c = 0; c < p.n_chunks; ++c) {
            // P of 8 rings x 4 steps, transposed into [M][item] (coalesced reads over M, conflict-free stride-33 writes)
            for (int item = warp; item < 32; item += 8) {
                const int r = item >> 2, t = item & 3, ring = 8 * c + r;
                if (ring < p.n_theta && t < nt) {
                    const double2* src = p.P + ((t0 + t) * p.n_theta + ring) * (int64_t)p.nm;
                    for (int m = lane; m < p.nm; m += 32) *reinterpret_cast<double2*>(sm + m * THETA_FSTRIDE + 2 * item) = src[m];
                }
            }
            double2 w[MAXT];
            {
                const double2* wf = reinterpret_cast<const double2*>(p.wtfrag + (int64_t)c * p.wt_stride) + lane;
#pragma unroll
                for (int s = 0; s < MAXT; ++s) w[s] = __ldg(wf + (warp + s * 8) * 32);
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < MAXT; ++s) {
                const double* bsrc = sm + tile_m[s] + fr;
                dmma_p(acc[s][0], acc[s][1], w[s].x, bsrc[0]);
                dmma_p(acc[s][0], acc[s][1], w[s].y, bsrc[32]);
            }
            __syncthreads();
        }
        const int t = lane & 3;
#pragma unroll
        for (int s = 0; s < MAXT; ++s) {
            const int2 tl = __ldg(p.tiles + warp + s * 8);
            const int l = tl.y + (lane >> 2), M = tl.x - p.L;
            if (l <= p.L && t < nt) p.out[(t0 + t) * p.n_out + l * (l + 1) + M - p.ell_min2] = make_double2(acc[s][0], acc[s][1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Cluster variant: a pair of CTAs (thread-block cluster of 2, two SMs) owns one block of 4 time steps.  CTA r stages only
// factor r's modes (76 KB instead of 152), so the F buffers fit twice and the stages overlap:
//   warps 0-7  ("A/C"): theta synthesis of the own factor for ring chunk c+1, written into the F buffers of BOTH CTAs
//                       (distributed shared memory), then the quadrature of chunk c for the own half of the M range;
//   warps 8-15 ("B")  : m-convolution of chunk c for the own half of the M range, P_M written over F1 of that buffer.
// One cluster barrier per chunk publishes the remote F writes and recycles the buffers, so the L2 table stream of stage A
// flies under the FMA work of stage B.  Measured (config-4 size, 2e4 steps): 9.3 ms against 7.4 ms for the single-CTA kernel -
// the stages do overlap, but each role now has 8 warps per SM instead of 16 and both are latency bound, which costs more than
// the overlap returns.  Kept as a selectable shape (product_tables(..., shape=2)); not the default.
struct ClusterParams {
    const double2* a[2];
    int n[2];
    const int* perm[2];
    double2* out;
    int64_t n_times;
    int n_out;
    const unsigned* ctl;    // [2][n_ctl_r]
    int n_ctl_r;
    const double* lamfrag;
    int64_t lam_stride;
    const int2* tiles;      // [2][tiles_r]
    int tiles_r;
    const double* wtfrag;   // [n_chunks][2][tiles_r * 64]
    int ell1, ell2, L_out, n_phi, n_chunks, qmax;
    int szA, bufStride, offF2rel, smem_doubles;
    int gbase[2];           // first global k-step of factor r (its modes sit at k-step g - gbase in shared memory)
    int g0[2], gcnt[2];     // groups of GM consecutive M convolved by CTA r
    int skip;
};

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int GM, int MAXT, int DA>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1)
modes_product_cluster_kernel(const ClusterParams p) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) double sm[];
    double* smr = cluster.map_shared_rank(sm, rank ^ 1);          // the partner CTA's shared memory
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, w8 = warp & 7;
    const bool isA = warp < 8;
    const int n_mout = 2 * p.L_out + 1;
    const int64_t n_tg = (p.n_times + PRODUCT_T - 1) / PRODUCT_T;
    const int fr = (lane & 3) * 8 + (lane >> 2);
    const double2* __restrict__ a = p.a[rank];
    const int n = p.n[rank];
    const int* __restrict__ perm = p.perm[rank];
    const int gbase = p.gbase[rank];

    for (int i = tid; i < p.smem_doubles / 2; i += 512) reinterpret_cast<double2*>(sm)[i] = make_double2(0.0, 0.0);
    unsigned* sctl = reinterpret_cast<unsigned*>(sm + p.smem_doubles);
    for (int i = tid; i < p.n_ctl_r; i += 512) sctl[i] = __ldg(p.ctl + rank * p.n_ctl_r + i);
    int tile_m[MAXT];
#pragma unroll
    for (int s = 0; s < MAXT; ++s) tile_m[s] = __ldg(p.tiles + rank * p.tiles_r + w8 + s * 8).x * 64;
    cluster.sync();   // both CTAs are initialised before the first remote write

    auto stageA = [&](int c, int bufoff) {
        const double* lf = p.lamfrag + (int64_t)c * p.lam_stride + lane;
        const int s0 = (int)sctl[w8], ns = (int)sctl[w8 + 1] - s0;
        const unsigned* ctl = sctl + s0;
        double fq[DA];
        unsigned cq[DA];
#pragma unroll
        for (int d = 0; d < DA; ++d) {
            cq[d] = ctl[d];
            fq[d] = __ldg(lf + (cq[d] & 0xffffu) * 32);
        }
        double c0[3] = {0.0, 0.0, 0.0}, c1[3] = {0.0, 0.0, 0.0};
        for (int base = 0; base < ns; base += DA) {
            ctl += DA;
#pragma unroll
            for (int d = 0; d < DA; ++d) {
                const unsigned u = cq[d];
                dmma_p(c0[d % 3], c1[d % 3], fq[d], sm[((int)(u & 0xffffu) - gbase) * 32 + fr]);
                if (u >> 31) {
                    const int off = bufoff + (int)((u >> 16) & 0x7fffu) * 64 + 2 * lane;
                    const double2 v = make_double2(c0[d % 3], c1[d % 3]);
                    *reinterpret_cast<double2*>(sm + off) = v;
                    *reinterpret_cast<double2*>(smr + off) = v;   // distributed shared memory: the partner convolves it too
                    c0[d % 3] = c1[d % 3] = 0.0;
                }
                cq[d] = ctl[d];
                fq[d] = __ldg(lf + (cq[d] & 0xffffu) * 32);
            }
        }
    };

    for (int64_t tg = blockIdx.x >> 1; tg < n_tg; tg += gridDim.x >> 1) {
        const int64_t t0 = tg * PRODUCT_T;
        const int nt = (int)min((int64_t)PRODUCT_T, p.n_times - t0);
        if (nt < PRODUCT_T) {   // ragged last block: clear the time slots that are not loaded
            for (int i = tid; i < p.szA / 2; i += 512) reinterpret_cast<double2*>(sm)[i] = make_double2(0.0, 0.0);
            __syncthreads();
        }
        for (int e = tid; e < nt * n; e += 512) {
            int t = e / n, idx = e - t * n;
            cp_async16_p(sm + __ldg(perm + idx) + 2 * t, a + t0 * n + e);
        }
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();

        double acc[MAXT][2];
#pragma unroll
        for (int s = 0; s < MAXT; ++s) acc[s][0] = acc[s][1] = 0.0;

        if (isA && !(p.skip & 1)) stageA(0, p.szA);
        cluster.sync();
        for (int c = 0; c < p.n_chunks; ++c) {
            const int bufoff = p.szA + (c & 1) * p.bufStride;
            if (isA) {
                if (c + 1 < p.n_chunks && !(p.skip & 1)) stageA(c + 1, p.szA + ((c + 1) & 1) * p.bufStride);
                double2 w[MAXT];
                {
                    const double2* wf = reinterpret_cast<const double2*>(p.wtfrag + ((int64_t)c * 2 + rank) * p.tiles_r * 64) + lane;
#pragma unroll
                    for (int s = 0; s < MAXT; ++s) w[s] = __ldg(wf + (w8 + s * 8) * 32);
                }
                named_bar_sync(1, 512);   // P_M of this chunk is in place
                if (!(p.skip & 4))
#pragma unroll
                for (int s = 0; s < MAXT; ++s) {
                    const double* bsrc = sm + bufoff + tile_m[s] + fr;
                    dmma_p(acc[s][0], acc[s][1], w[s].x, bsrc[0]);
                    dmma_p(acc[s][0], acc[s][1], w[s].y, bsrc[32]);
                }
            } else {
                double2 pacc[GM];
#pragma unroll
                for (int q = 0; q < GM; ++q) pacc[q] = make_double2(0.0, 0.0);
                const int gq = (int)sctl[9 + w8];   // my group of GM consecutive M, -1: none
                const bool mine = gq >= 0;
                if (mine && !(p.skip & 2)) {
                    const int M0 = -p.L_out + GM * gq;
                    const double2* f1 = reinterpret_cast<const double2*>(sm + bufoff) + lane;
                    const double2* f2 = reinterpret_cast<const double2*>(sm + bufoff + p.offF2rel) + lane;
                    const int l2 = p.ell2;
                    for (int q = -p.qmax; q <= p.qmax; ++q) {
                        const int Me = M0 + q * p.n_phi;
                        const int lo = max(-p.ell1, Me - l2), hi = min(p.ell1, Me + GM - 1 + l2);
                        if (lo > hi) continue;
                        double2 y[GM + 3];
                        const double2* f2q = f2 + (Me + l2) * 32;
#pragma unroll
                        for (int k = 0; k < GM - 1; ++k) y[k] = f2q[(1 + k - lo) * 32];
                        auto block = [&](int m1b, auto tail) {
#pragma unroll
                            for (int k = GM - 2; k >= 0; --k) y[k + 4] = y[k];
#pragma unroll
                            for (int k = 0; k < 4; ++k) y[k] = f2q[(k - 3 - m1b) * 32];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                double2 x = make_double2(0.0, 0.0);
                                if (!decltype(tail)::value || m1b + i <= hi) x = f1[(m1b + i + p.ell1) * 32];
#pragma unroll
                                for (int qq = 0; qq < GM; ++qq) cfma(pacc[qq], x, y[qq - i + 3]);
                            }
                        };
                        int m1b = lo;
#pragma unroll 2
                        for (; m1b + 3 <= hi; m1b += 4) block(m1b, std::false_type{});
                        if (m1b <= hi) block(m1b, std::true_type{});
                    }
                }
                named_bar_sync(2, 256);   // every convolution warp is done reading F1 of this buffer
                if (mine) {
                    double2* sP = reinterpret_cast<double2*>(sm + bufoff);
#pragma unroll
                    for (int qq = 0; qq < GM; ++qq)
                        if (GM * gq + qq < n_mout) sP[(GM * gq + qq) * 32 + lane] = pacc[qq];
                }
                named_bar_sync(1, 512);
            }
            cluster.sync();   // remote F writes of chunk c+1 are visible; the buffer of chunk c may be refilled
        }

        if (isA) {
            const int t = lane & 3;
#pragma unroll
            for (int s = 0; s < MAXT; ++s) {
                const int2 tl = __ldg(p.tiles + rank * p.tiles_r + w8 + s * 8);
                const int l = tl.y + (lane >> 2), M = tl.x - p.L_out;
                if (l <= p.L_out && t < nt) p.out[(t0 + t) * p.n_out + l * (l + 1) + M] = make_double2(acc[s][0], acc[s][1]);
            }
        }
    }
    cluster.sync();   // no CTA exits while its partner may still write into its shared memory
}

}  // namespace scrib200

extern "C" size_t scrib200_modes_product_max_shared_bytes(void) { return 227u * 1024u; }

extern "C" int scrib200_theta_quad(const double* P, int64_t n_times, const int* tiles, int n_tiles, const double* wtfrag,
                                   int64_t wt_stride, const int* cfg, double* out, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(P && tiles && wtfrag && cfg && out, "theta_quad: null pointer");
    SCRIB200_REQUIRE(aligned16(P) && aligned16(out) && aligned16(wtfrag), "theta_quad: pointers must be 16-byte aligned");
    QuadParams p;
    p.P = reinterpret_cast<const double2*>(P);
    p.out = reinterpret_cast<double2*>(out);
    p.n_times = n_times;
    p.n_theta = cfg[0];
    p.nm = cfg[1];
    p.n_chunks = cfg[2];
    p.n_out = cfg[3];
    p.ell_min2 = cfg[4];
    p.L = cfg[5];
    p.tiles = reinterpret_cast<const int2*>(tiles);
    p.wtfrag = wtfrag;
    p.wt_stride = wt_stride;
    SCRIB200_REQUIRE(n_tiles == 8 * 21 && wt_stride == (int64_t)n_tiles * 64, "theta_quad: the tile table must hold 8 x 21 tiles (l_max up to ~32)");
    SCRIB200_REQUIRE(p.n_theta > 0 && p.nm == 2 * p.L + 1 && p.n_chunks > 0 && p.n_out > 0, "theta_quad: bad table configuration");
    const size_t smem = (size_t)p.nm * THETA_FSTRIDE * sizeof(double);
    if (n_times <= 0) return SCRIB200_OK;
    const int64_t n_tg = (n_times + PRODUCT_T - 1) / PRODUCT_T;
    const int64_t grid = n_tg < 296 ? n_tg : 296;
    auto kern = theta_quad_kernel<21>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("theta_quad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return SCRIB200_ECUDA;
        }
    }
    kern<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(p);
    SCRIB200_CHECK_LAUNCH("theta_quad");
    return SCRIB200_OK;
}

extern "C" int scrib200_theta_synth(const double* modes, int n_modes, int64_t n_times, const int* perm, const int* ctl, int n_ctl,
                                    const double* lamfrag, int64_t lam_stride, const int* cfg, double* out, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(modes && perm && ctl && lamfrag && cfg && out, "theta_synth: null pointer");
    SCRIB200_REQUIRE(aligned16(modes) && aligned16(out), "theta_synth: pointers must be 16-byte aligned");
    ThetaParams p;
    p.a = reinterpret_cast<const double2*>(modes);
    p.n = n_modes;
    p.perm = perm;
    p.ctl = reinterpret_cast<const unsigned*>(ctl);
    p.n_ctl = n_ctl;
    p.lamfrag = lamfrag;
    p.lam_stride = lam_stride;
    p.out = reinterpret_cast<double2*>(out);
    p.n_times = n_times;
    p.n_theta = cfg[0];
    p.nm = cfg[1];
    p.n_chunks = cfg[2];
    p.szA = cfg[3];
    p.smem_doubles = cfg[4];
    SCRIB200_REQUIRE(n_modes > 0 && n_ctl > 9 && p.n_theta > 0 && p.nm > 0 && p.n_chunks > 0 && (p.szA & 1) == 0 && (p.smem_doubles & 1) == 0,
                     "theta_synth: bad table configuration");
    SCRIB200_REQUIRE(p.smem_doubles >= p.szA + p.nm * THETA_FSTRIDE, "theta_synth: shared-memory layout too small for %d values of m", p.nm);
    const size_t smem = (size_t)p.smem_doubles * sizeof(double) + (size_t)n_ctl * sizeof(unsigned);
    SCRIB200_REQUIRE(smem <= 113u * 1024u, "theta_synth: %zu bytes of shared memory needed (limit %u, two CTAs per SM); use the dense synthesis", smem, 113u * 1024u);
    if (n_times <= 0) return SCRIB200_OK;
    const int64_t n_tg = (n_times + PRODUCT_T - 1) / PRODUCT_T;
    const int64_t grid = n_tg < 296 ? n_tg : 296;
    auto kern = theta_synth_kernel<9>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    if (e != cudaSuccess) {
        set_error("theta_synth: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return SCRIB200_ECUDA;
    }
    kern<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(p);
    SCRIB200_CHECK_LAUNCH("theta_synth");
    return SCRIB200_OK;
}

extern "C" int scrib200_modes_product(const double* a1, int n1, const double* a2, int n2, int64_t n_times,
                                      const int* perm1, const int* perm2, const int* ctl, int n_steps,
                                      const double* lamfrag, int64_t lam_stride, const int* tiles, int n_tiles,
                                      const double* wtfrag, int64_t wt_stride, const int* cfg, double* out,
                                      int n_ctas, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(a1 && a2 && perm1 && perm2 && ctl && lamfrag && tiles && wtfrag && cfg && out, "modes_product: null pointer");
    SCRIB200_REQUIRE(aligned16(a1) && aligned16(a2) && aligned16(out) && aligned16(wtfrag), "modes_product: pointers must be 16-byte aligned");
    if (cfg[15] == 1) {   // cluster variant: per-CTA-rank tables (scri_b200/_product.py:product_tables(..., shape=2))
        ClusterParams c;
        c.a[0] = reinterpret_cast<const double2*>(a1);
        c.a[1] = reinterpret_cast<const double2*>(a2);
        c.n[0] = n1;
        c.n[1] = n2;
        c.perm[0] = perm1;
        c.perm[1] = perm2;
        c.out = reinterpret_cast<double2*>(out);
        c.n_times = n_times;
        c.ctl = reinterpret_cast<const unsigned*>(ctl);
        c.n_ctl_r = n_steps;
        c.lamfrag = lamfrag;
        c.lam_stride = lam_stride;
        c.tiles = reinterpret_cast<const int2*>(tiles);
        c.tiles_r = n_tiles;
        c.wtfrag = wtfrag;
        c.ell1 = cfg[0];
        c.ell2 = cfg[1];
        c.L_out = cfg[2];
        c.n_phi = cfg[3];
        c.n_chunks = cfg[4];
        c.qmax = cfg[5];
        c.szA = cfg[6];
        c.bufStride = cfg[7];
        c.offF2rel = cfg[8];
        c.smem_doubles = cfg[9];
        c.skip = cfg[14];
        c.gbase[0] = 0;
        c.gbase[1] = cfg[16];
        c.g0[0] = cfg[17];
        c.gcnt[0] = cfg[18];
        c.g0[1] = cfg[19];
        c.gcnt[1] = cfg[20];
        c.n_out = (c.L_out + 1) * (c.L_out + 1);
        SCRIB200_REQUIRE(cfg[10] == 16 && cfg[12] == 5 && cfg[13] == 12 && n_tiles == 8 * 12, "modes_product: unsupported cluster kernel shape");
        SCRIB200_REQUIRE(c.gcnt[0] >= 0 && c.gcnt[0] <= 8 && c.gcnt[1] >= 0 && c.gcnt[1] <= 8, "modes_product: more than 8 groups of M per CTA");
        SCRIB200_REQUIRE(wt_stride == (int64_t)2 * n_tiles * 64, "modes_product: cluster quadrature table stride");
        SCRIB200_REQUIRE((c.szA & 1) == 0 && (c.bufStride & 1) == 0 && (c.offF2rel & 1) == 0 && (c.smem_doubles & 1) == 0, "modes_product: shared-memory offsets must be even");
        const size_t smem = (size_t)c.smem_doubles * sizeof(double) + (size_t)n_steps * sizeof(unsigned);
        SCRIB200_REQUIRE(smem <= scrib200_modes_product_max_shared_bytes(), "modes_product: %zu bytes of shared memory needed (limit %zu)", smem,
                         scrib200_modes_product_max_shared_bytes());
        if (n_times <= 0) return SCRIB200_OK;
        const int64_t n_tg = (n_times + PRODUCT_T - 1) / PRODUCT_T;
        if (n_ctas <= 0) n_ctas = 148;
        int64_t pairs = n_ctas / 2;
        if (pairs > n_tg) pairs = n_tg;
        if (pairs < 1) pairs = 1;
        auto kern = modes_product_cluster_kernel<5, 12, 9>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_error("modes_product: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return SCRIB200_ECUDA;
        }
        kern<<<(unsigned)(2 * pairs), 512, smem, (cudaStream_t)stream>>>(c);
        SCRIB200_CHECK_LAUNCH("modes_product(cluster)");
        return SCRIB200_OK;
    }
    ProductParams p;
    p.a1 = reinterpret_cast<const double2*>(a1);
    p.a2 = reinterpret_cast<const double2*>(a2);
    p.out = reinterpret_cast<double2*>(out);
    p.n_times = n_times;
    p.n1 = n1;
    p.n2 = n2;
    p.perm1 = perm1;
    p.perm2 = perm2;
    p.ctl = reinterpret_cast<const unsigned*>(ctl);
    p.n_steps = n_steps;
    p.lamfrag = lamfrag;
    p.lam_stride = lam_stride;
    p.tiles = reinterpret_cast<const int2*>(tiles);
    p.n_tiles = n_tiles;
    p.wtfrag = wtfrag;
    p.wt_stride = wt_stride;
    p.ell1 = cfg[0];
    p.ell2 = cfg[1];
    p.L_out = cfg[2];
    p.n_phi = cfg[3];
    p.n_chunks = cfg[4];
    p.qmax = cfg[5];
    p.szA = cfg[6];
    p.offF1 = cfg[7];
    p.offF2 = cfg[8];
    const int smem_doubles = cfg[9], nwarps = cfg[10], gm = cfg[12], maxt = cfg[13];
    p.skip = cfg[14];
    p.smem_doubles = smem_doubles;
    p.n_out = (p.L_out + 1) * (p.L_out + 1);
    p.stage_out = (PRODUCT_T * p.n_out * 2 <= p.szA) ? 1 : 0;
    SCRIB200_REQUIRE(n1 > 0 && n2 > 0 && n_steps > 0 && n_steps < 65536 && n_tiles > 0 && p.n_chunks > 0, "modes_product: empty tables");
    SCRIB200_REQUIRE(p.ell1 >= 0 && p.ell2 >= 0 && p.L_out >= 0 && p.n_phi > 0 && p.qmax >= 0, "modes_product: bad band limits");
    SCRIB200_REQUIRE((p.szA & 1) == 0 && (p.offF1 & 1) == 0 && (p.offF2 & 1) == 0, "modes_product: shared-memory offsets must be even");
    const size_t smem = (size_t)smem_doubles * sizeof(double) + (size_t)n_steps * sizeof(unsigned);
    SCRIB200_REQUIRE(smem <= scrib200_modes_product_max_shared_bytes(), "modes_product: %zu bytes of shared memory needed (limit %zu); use the dense path", smem,
                     scrib200_modes_product_max_shared_bytes());
    SCRIB200_REQUIRE((gm == 5 && nwarps == 16 && maxt == 11) || (gm == 9 && nwarps == 8 && maxt == 21),
                     "modes_product: unsupported kernel shape (gm %d, warps %d, tiles per warp %d)", gm, nwarps, maxt);
    const int n_groups = (2 * p.L_out + 1 + gm - 1) / gm;
    SCRIB200_REQUIRE(nwarps >= n_groups, "modes_product: %d warps for %d groups of M", nwarps, n_groups);
    SCRIB200_REQUIRE(n_tiles == nwarps * maxt, "modes_product: the tile table must be padded to warps x tiles per warp (%d != %d x %d)", n_tiles, nwarps, maxt);
    if (n_times <= 0) return SCRIB200_OK;
    int64_t n_tg = (n_times + PRODUCT_T - 1) / PRODUCT_T;
    if (n_ctas <= 0) n_ctas = 148;
    int64_t grid = n_tg < n_ctas ? n_tg : n_ctas;
    auto kern = (gm == 5) ? modes_product_kernel<5, 11, 9, 512> : modes_product_kernel<9, 21, 18, 256>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
        set_error("modes_product: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return SCRIB200_ECUDA;
    }
    kern<<<(unsigned)grid, nwarps * 32, smem, (cudaStream_t)stream>>>(p);
    SCRIB200_CHECK_LAUNCH("modes_product");
    return SCRIB200_OK;
}
