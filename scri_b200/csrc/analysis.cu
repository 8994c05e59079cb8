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
