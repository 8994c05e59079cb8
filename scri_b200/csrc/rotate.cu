// K4: per-time-step Wigner-D rotation of the modes, in place.
//
// Replaces scri/rotations.py:346-392 (numba loops) + sf._Wigner_D_matrices (per time step).
// a'_{l m}(t) = sum_{m'} a_{l m'}(t) D^l_{m',m}(R(t)).
//
// D is never materialised: D^l_{m'm} = PhA(m'+m) PhB(m-m') P^l_{m'm}(cos beta), with the phases taken
// from per-step tables of powers of Ra, Rb (no trigonometry, regular at the poles) and the real
// polynomial P advanced in l by the three-term recurrence from an exact seed at l0 = max(|m'|,|m|)
// (tables built once on the host: scri_b200/_sf.py:wigner_tables).
//
// Mapping: a CTA owns TB consecutive time steps; the [TB, n_modes] tile is staged through shared
// memory with coalesced loads/stores; one thread per (time step, output m) walks m' and l.
// HBM-bound by design: 2*16*n_modes + 32 bytes per time step.
#include "common.cuh"

namespace scrib200 {

__global__ void rotate_modes_kernel(double2* __restrict__ data, int64_t n_times, int ell_min, int ell_max,
                                    const double2* __restrict__ spinors, int64_t spinor_stride,
                                    const double* __restrict__ seed, const double* __restrict__ rec, int TB) {
    extern __shared__ double2 sm[];
    const int L = ell_max;
    const int nm = 2 * L + 1;
    const int n_modes = L * (L + 2) - ell_min * ell_min + 1;
    double2* s_in = sm;                       // [TB][n_modes]
    double2* s_out = s_in + TB * n_modes;     // [TB][n_modes]
    double2* s_pa = s_out + TB * n_modes;     // [TB][nm]   Ra^k, k = 0..2L
    double2* s_pb = s_pa + TB * nm;           // [TB][nm]   Rb^k
    double* s_cos = reinterpret_cast<double*>(s_pb + TB * nm);  // [TB]

    const int64_t t0 = (int64_t)blockIdx.x * TB;
    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
    const int tile = TB * n_modes;

    for (int idx = tid; idx < tile; idx += nthreads) {
        int64_t t = t0 + idx / n_modes;
        s_in[idx] = (t < n_times) ? data[t0 * n_modes + idx] : make_double2(0.0, 0.0);
        s_out[idx] = make_double2(0.0, 0.0);
    }
    // powers of the (normalised) spinor components
    for (int idx = tid; idx < 2 * TB; idx += nthreads) {
        int tt = idx >> 1, which = idx & 1;
        int64_t t = t0 + tt;
        if (t >= n_times) t = n_times - 1;
        double2 Ra = spinors[t * spinor_stride + 0];
        double2 Rb = spinors[t * spinor_stride + 1];
        double ra2 = Ra.x * Ra.x + Ra.y * Ra.y;
        double rb2 = Rb.x * Rb.x + Rb.y * Rb.y;
        double n2 = ra2 + rb2;
        double inv = 1.0 / sqrt(n2);
        double2 z = which ? cscale(inv, Rb) : cscale(inv, Ra);
        double2* p = (which ? s_pb : s_pa) + tt * nm;
        double2 acc = make_double2(1.0, 0.0);
        p[0] = acc;
        for (int k = 1; k < nm; ++k) {
            acc = cmul(acc, z);
            p[k] = acc;
        }
        // rb2 == 0 exactly: D is diagonal (Ra^{2m}); flag it by cos = 2 so an identity rotation is bit-exact
        if (!which) s_cos[tt] = (rb2 == 0.0) ? 2.0 : (ra2 - rb2) / n2;
    }
    __syncthreads();

    for (int idx = tid; idx < TB * nm; idx += nthreads) {
        const int tt = idx / nm;
        const int mi = idx - tt * nm;
        const int m = mi - L;
        const int am = m < 0 ? -m : m;
        if (t0 + tt >= n_times) continue;
        const double cosb = s_cos[tt];
        const bool diagonal = cosb > 1.5;
        const double2* pa = s_pa + tt * nm;
        const double2* pb = s_pb + tt * nm;
        const double2* in = s_in + tt * n_modes;
        double2* out = s_out + tt * n_modes;
        const int lmin2 = ell_min * ell_min;
        for (int mp = -L; mp <= L; ++mp) {
            if (diagonal && mp != m) continue;
            const int amp = mp < 0 ? -mp : mp;
            const int l0 = amp > am ? amp : am;
            const int ka = mp + m, kb = m - mp;
            double2 fa = ka >= 0 ? pa[ka] : cconj(pa[-ka]);
            double2 fb = kb >= 0 ? pb[kb] : cconj(pb[-kb]);
            const double2 ph = cmul(fa, fb);
            double P = diagonal ? 1.0 : seed[(mp + L) * nm + mi];
            double Pm1 = 0.0;
            const double* rc = rec + ((size_t)(mp + L) * nm + mi) * 3;
            const size_t rstride = (size_t)nm * nm * 3;
            for (int l = l0; l <= L; ++l) {
                if (l >= ell_min) {
                    const int base = l * (l + 1) - lmin2;
                    double2 w = cscale(P, ph);
                    cfma(out[base + m], in[base + mp], w);
                }
                if (l < L && !diagonal) {
                    const double* c3 = rc + rstride * l;
                    double Pn = (c3[0] * cosb - c3[1]) * P - c3[2] * Pm1;
                    Pm1 = P;
                    P = Pn;
                }
            }
        }
    }
    __syncthreads();

    for (int idx = tid; idx < tile; idx += nthreads) {
        int64_t t = t0 + idx / n_modes;
        if (t < n_times) data[t0 * n_modes + idx] = s_out[idx];
    }
}

}  // namespace scrib200

extern "C" int scrib200_rotate_modes(double* data, int64_t n_times, int ell_min, int ell_max, const double* spinors,
                                     int64_t spinor_stride, const double* seed, const double* rec, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(data && spinors && seed && rec, "rotate_modes: null pointer");
    SCRIB200_REQUIRE(ell_min >= 0 && ell_max >= ell_min, "rotate_modes: bad ell range [%d, %d]", ell_min, ell_max);
    SCRIB200_REQUIRE(spinor_stride == 0 || spinor_stride == 2, "rotate_modes: spinor_stride must be 0 or 2");
    SCRIB200_REQUIRE(aligned16(data) && aligned16(spinors), "rotate_modes: pointers must be 16-byte aligned");
    if (n_times <= 0) return SCRIB200_OK;
    const int L = ell_max, nm = 2 * L + 1;
    const int n_modes = L * (L + 2) - ell_min * ell_min + 1;
    int TB = 256 / nm;
    if (TB < 1) TB = 1;
    size_t smem = (size_t)TB * (2 * n_modes + 2 * nm) * sizeof(double2) + TB * sizeof(double);
    while (smem > 200 * 1024 && TB > 1) {
        --TB;
        smem = (size_t)TB * (2 * n_modes + 2 * nm) * sizeof(double2) + TB * sizeof(double);
    }
    SCRIB200_REQUIRE(smem <= 227 * 1024, "rotate_modes: ell_max=%d needs %zu bytes of shared memory", L, smem);
    int threads = TB * nm;
    threads = ((threads + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    cudaFuncSetAttribute(rotate_modes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (n_times + TB - 1) / TB;
    rotate_modes_kernel<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<double2*>(data), n_times, ell_min, ell_max, reinterpret_cast<const double2*>(spinors),
        spinor_stride, seed, rec, TB);
    SCRIB200_CHECK_LAUNCH("rotate_modes");
    return SCRIB200_OK;
}
