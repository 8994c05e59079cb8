// K10: Weyl-scalar mixing under a BMS transformation (elementwise over the synthesized grid).
//
// Replaces scri/waveform_grid.py:504-550, 559 (psi0..psi3 pick up the higher Weyl scalars multiplied by powers of the
// time-dependent factor  z = eth u' / k = (t - alpha) gamma k eth(v.r) - eth alpha) and the Horner ladders of
// scri/asymptotic_bondi_data/transformations.py:340-390 (z = (eth k / k)(u - alpha) - eth alpha).  Both are
//     out[i, g] = scale[g] * ( c_0 F_0[i,g] + c_1 F_1[i,g] z + ... + c_Q F_Q[i,g] z^Q - offset[g] ),
//     z[i, g]   = (t[i] - alpha[g]) * A[g] - C[g]                                    (complex A, C; Q <= 4)
// evaluated by Horner's rule from the highest field.  HBM-bound: (Q+1) reads + 1 write of 16 bytes per element; one
// thread owns one grid point (its constants stay in registers) and walks a slab of time steps, lanes along g.
#include "common.cuh"

namespace scrib200 {

struct MixFields {
    const double2* F[5];
    double c[5];
};

__global__ void __launch_bounds__(128)
weyl_mix_kernel(MixFields f, int nq, const double* __restrict__ t, int64_t N, int G, const double* __restrict__ alpha,
                const double2* __restrict__ A, const double2* __restrict__ C, const double* __restrict__ scale,
                const double2* __restrict__ offset, double2* __restrict__ out, int rows_per_cta) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const double al = alpha[g], sc = scale[g];
    const double2 a = A[g], c = C[g];
    const double2 off = offset ? offset[g] : make_double2(0.0, 0.0);
    const int64_t i0 = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t i1 = (i0 + rows_per_cta < N) ? i0 + rows_per_cta : N;
    for (int64_t i = i0; i < i1; ++i) {
        const double dt = t[i] - al;
        const double2 z = make_double2(dt * a.x - c.x, dt * a.y - c.y);
        const int64_t e = i * G + g;
        double2 acc = cscale(f.c[nq - 1], f.F[nq - 1][e]);
        for (int q = nq - 2; q >= 0; --q) {
            acc = cmul(acc, z);
            const double2 v = f.F[q][e];
            acc.x = fma(f.c[q], v.x, acc.x);
            acc.y = fma(f.c[q], v.y, acc.y);
        }
        out[e] = make_double2(sc * (acc.x - off.x), sc * (acc.y - off.y));
    }
}

// out[e] = a[e] * b[e] (complex), the pointwise product of ModesTimeSeries.grid_multiply (modes_time_series.py:190).
__global__ void __launch_bounds__(256)
grid_product_kernel(const double2* __restrict__ a, const double2* __restrict__ b, double2* __restrict__ out, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = cmul(a[e], b[e]);
}

}  // namespace scrib200

extern "C" int scrib200_grid_product(const double* a, const double* b, double* out, int64_t n, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(a && b && out, "grid_product: null pointer");
    SCRIB200_REQUIRE(aligned16(a) && aligned16(b) && aligned16(out), "grid_product: pointers must be 16-byte aligned");
    if (n <= 0) return SCRIB200_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    grid_product_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2*>(a), reinterpret_cast<const double2*>(b), reinterpret_cast<double2*>(out), n);
    SCRIB200_CHECK_LAUNCH("grid_product");
    return SCRIB200_OK;
}

extern "C" int scrib200_weyl_mix(const double* const* fields, const double* coef, int n_fields, const double* t,
                                 int64_t n_times, int G, const double* alpha, const double* A, const double* C,
                                 const double* scale, const double* offset, double* out, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(fields && coef && t && alpha && A && C && scale && out, "weyl_mix: null pointer");
    SCRIB200_REQUIRE(n_fields >= 1 && n_fields <= 5, "weyl_mix: n_fields=%d must be 1..5", n_fields);
    SCRIB200_REQUIRE(G > 0, "weyl_mix: G=%d", G);
    if (n_times <= 0) return SCRIB200_OK;
    MixFields f;
    for (int q = 0; q < 5; ++q) {
        f.F[q] = (q < n_fields) ? reinterpret_cast<const double2*>(fields[q]) : nullptr;
        f.c[q] = (q < n_fields) ? coef[q] : 0.0;
        SCRIB200_REQUIRE(q >= n_fields || (fields[q] && aligned16(fields[q])), "weyl_mix: field %d is null or misaligned", q);
    }
    SCRIB200_REQUIRE(aligned16(A) && aligned16(C) && aligned16(out) && aligned16(offset), "weyl_mix: pointers must be 16-byte aligned");
    const int rows = 64;
    dim3 grid((G + 127) / 128, (unsigned)((n_times + rows - 1) / rows));
    SCRIB200_REQUIRE(grid.y <= 65535u * 16u, "weyl_mix: series too long");
    weyl_mix_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
        f, n_fields, t, n_times, G, alpha, reinterpret_cast<const double2*>(A), reinterpret_cast<const double2*>(C), scale,
        reinterpret_cast<const double2*>(offset), reinterpret_cast<double2*>(out), rows);
    SCRIB200_CHECK_LAUNCH("weyl_mix");
    return SCRIB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// The packed synthesis operand of scrib200_swsh_synthesize, built on the device from the rotor grid:
//   Y[g, (l, m)] = (-1)^s sqrt((2l+1)/4pi) D^l_{m,-s}(R_g)      (sf.SWSH_grid, scri/waveform_grid.py:470-471)
//   B[2lm, 2g] = Re Y, B[2lm+1, 2g] = -Im Y, B[2lm, 2g+1] = Im Y, B[2lm+1, 2g+1] = Re Y   (lm counted from ell_min).
// One thread per (grid point, m): phases from powers of the spinor components, the real factor advanced in l by the
// three-term recurrence (seed / factored coefficient tables of scri_b200._sf, the same as the rotation kernel's).
namespace scrib200 {

__global__ void __launch_bounds__(128)
swsh_pack_kernel(const double* __restrict__ R, int G, int s, int ell_min, int ell_max, const double* __restrict__ seed,
                 const double2* __restrict__ uv, int Lt, double* __restrict__ B, int Ncpad) {
    const int nm_out = 2 * ell_max + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= G * nm_out) return;
    const int g = idx / nm_out, mp = idx - g * nm_out - ell_max;
    const int m = -s;                                  // column of D
    const int nmt = 2 * Lt + 1;
    const double w = R[4 * g], x = R[4 * g + 1], y = R[4 * g + 2], z = R[4 * g + 3];
    const double2 Ra = make_double2(w, z), Rb = make_double2(y, x);
    const double cosb = (Ra.x * Ra.x + Ra.y * Ra.y) - (Rb.x * Rb.x + Rb.y * Rb.y);
    const int amp = mp < 0 ? -mp : mp, am = m < 0 ? -m : m;
    const int l0 = amp > am ? amp : am;
    if (l0 > ell_max) return;
    auto cpow = [](double2 zz, int k) {               // zz^k for k >= 0, conj(zz)^|k| for k < 0
        if (k < 0) { zz = cconj(zz); k = -k; }
        double2 acc = make_double2(1.0, 0.0);
        for (int i = 0; i < k; ++i) acc = cmul(acc, zz);
        return acc;
    };
    const double2 ph = cmul(cpow(Ra, mp + m), cpow(Rb, m - mp));
    double P = seed[(mp + Lt) * nmt + (m + Lt)], Pm1 = 0.0;
    const double sgn = (s & 1) ? -1.0 : 1.0;
    const double mmp = (double)(m * mp);
    for (int l = l0; l <= ell_max; ++l) {
        if (l >= ell_min) {
            const double f = sgn * sqrt((2.0 * l + 1.0) / (4.0 * 3.14159265358979323846)) * P;
            const double re = f * ph.x, im = f * ph.y;
            const size_t lm = (size_t)(l * (l + 1) - ell_min * ell_min + mp);
            double* b0 = B + (2 * lm) * Ncpad + 2 * g;
            double* b1 = b0 + Ncpad;
            b0[0] = re;
            b0[1] = im;
            b1[0] = -im;
            b1[1] = re;
        }
        if (l < ell_max) {
            const double2 f1 = uv[l * nmt + mp + Lt], f2 = uv[l * nmt + m + Lt];
            const double rl = (l > 0) ? 1.0 / (double)(l * (l + 1)) : 0.0;
            const double Pn = (f1.x * f2.x) * (cosb - mmp * rl) * P - (f1.y * f2.y) * Pm1;
            Pm1 = P;
            P = Pn;
        }
    }
}

}  // namespace scrib200

extern "C" int scrib200_swsh_pack(const double* rotors, int G, int spin, int ell_min, int ell_max, const double* seed,
                                  const double* uv, int table_ell_max, double* Bmat, int Kpad, int Ncpad, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(rotors && seed && uv && Bmat, "swsh_pack: null pointer");
    SCRIB200_REQUIRE(G > 0 && ell_min >= 0 && ell_max >= ell_min, "swsh_pack: bad sizes G=%d ell=[%d, %d]", G, ell_min, ell_max);
    const int as = spin < 0 ? -spin : spin;
    SCRIB200_REQUIRE(table_ell_max >= ell_max && table_ell_max >= as, "swsh_pack: tables for ell <= %d cannot serve ell_max=%d, |s|=%d",
                     table_ell_max, ell_max, as);
    const int n = ell_max * (ell_max + 2) - ell_min * ell_min + 1;
    SCRIB200_REQUIRE(Kpad >= 2 * n && Ncpad >= 2 * G, "swsh_pack: Bmat [%d, %d] too small for %d modes x %d points", Kpad, Ncpad, n, G);
    SCRIB200_REQUIRE(aligned16(uv), "swsh_pack: uv must be 16-byte aligned");
    cudaError_t e = cudaMemsetAsync(Bmat, 0, (size_t)Kpad * Ncpad * sizeof(double), (cudaStream_t)stream);
    SCRIB200_REQUIRE(e == cudaSuccess, "swsh_pack: %s", cudaGetErrorString(e));
    const int total = G * (2 * ell_max + 1);
    swsh_pack_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rotors, G, spin, ell_min, ell_max, seed,
                                                                             reinterpret_cast<const double2*>(uv), table_ell_max,
                                                                             Bmat, Ncpad);
    SCRIB200_CHECK_LAUNCH("swsh_pack");
    return SCRIB200_OK;
}
