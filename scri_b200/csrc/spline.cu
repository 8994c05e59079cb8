// K2: BMS retarded-time remap = batched not-a-knot cubic spline build + evaluation.
//
// Replaces scri/waveform_grid.py:576-588: for every grid point g the reference builds two scipy
// InterpolatedUnivariateSpline objects (k=3 interpolating spline, FITPACK not-a-knot) on the knots
//     x_i = k[g] * (t[i] - alpha[g])
// for Re and Im of F[:, g] and evaluates them at the common output times u'.
//
// The interpolating cubic spline is unique, so it is computed here in the moment (second-derivative)
// form: tridiagonal system  h_{i-1} M_{i-1} + 2(h_{i-1}+h_i) M_i + h_i M_{i+1} = 6(d_i - d_{i-1})
// with the not-a-knot conditions eliminated into the first and last rows, solved by a Thomas sweep.
// The knots are the reference's own rounded abscissae (x_i computed as mul(k, sub(t, alpha)) with no
// FMA contraction), because at t ~ 1e4 one ulp of x already moves the interpolant at the 1e-13 level.
//
// Parallelisation: one thread per (grid point, chunk of knots); lanes run along g so every load and
// store of the [time, g] arrays is coalesced.  A chunk solves its knots plus a halo of HALO knots on
// each side with a natural cut (M = 0); the tridiagonal inverse decays at least as fast as 2^-k
// (0.268^k on uniform knots), so HALO = 64 puts the cut below 1e-19 relative; chunks that reach a true
// end of the series apply the exact not-a-knot rows there.  The evaluation is fused into the back
// substitution (M_i and M_{i+1} are live exactly when interval i is visited), so the moments are never
// stored; the forward-sweep coefficients go through a coalesced workspace.
#include "common.cuh"

namespace scrib200 {

constexpr int HALO = 64;
constexpr int DEFAULT_CHUNK = 512;

struct Knots {
    const double* t;
    double k, alpha;
    __device__ __forceinline__ double x(int64_t i) const { return __dmul_rn(k, __dsub_rn(t[i], alpha)); }
};

__device__ __forceinline__ double2 eval_piece(double u, double xi, double xi1, double h, double inv_h, double2 yi,
                                              double2 yi1, double2 Mi, double2 Mi1) {
    const double A = (xi1 - u) * inv_h;
    const double B = (u - xi) * inv_h;
    const double h26 = h * h * (1.0 / 6.0);
    const double ca = (A * A * A - A) * h26;
    const double cb = (B * B * B - B) * h26;
    double2 r;
    r.x = A * yi.x + B * yi1.x + (ca * Mi.x + cb * Mi1.x);
    r.y = A * yi.y + B * yi1.y + (ca * Mi.y + cb * Mi1.y);
    return r;
}

// MODE 0: evaluate at up[] (the BMS remap); MODE 1 / 2: first / second derivative at the knots themselves
// (out is then [N, G]; `up`/`Nout` unused) - the CubicSpline(t, data).derivative(k)(t) of waveform_base.py:689-695.
template <int MODE>
__global__ void __launch_bounds__(64)
bms_spline_remap_kernel(const double* __restrict__ t, int64_t N, const double2* __restrict__ F, int G,
                        const double* __restrict__ kconf, const double* __restrict__ alpha,
                        const double* __restrict__ up, int64_t Nout, double2* __restrict__ out, int C, int W,
                        double* __restrict__ ws_c, double2* __restrict__ ws_d) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int64_t chunk = blockIdx.y;
    const int64_t a = chunk * C;                             // first interval of this chunk
    const int64_t b = (a + C < N - 1) ? a + C : N - 1;       // intervals a .. b-1, knots a .. b
    const bool last_chunk = (b == N - 1);
    const int64_t lo = (a - HALO > 0) ? a - HALO : 0;
    const int64_t hi = (b + HALO < N - 1) ? b + HALO : N - 1;
    const bool true_lo = (lo == 0), true_hi = (hi == N - 1);
    const int64_t rs = lo + 1, re = hi - 1;                  // rows solved (N >= 4 => rs <= re)

    Knots kn{t, kconf[g], alpha[g]};
    double* wc = ws_c + ((size_t)chunk * W) * G + g;         // slot s -> wc[s*G]
    double2* wd = ws_d + ((size_t)chunk * W) * G + g;
    const double2* Fg = F + g;

    // ---------------- forward elimination over rows rs..re
    double x_im1 = kn.x(rs - 1), x_i = kn.x(rs);
    double2 y_im1 = Fg[(rs - 1) * G], y_i = Fg[rs * G];
    double h_im1 = x_i - x_im1;
    double inv_him1 = 1.0 / h_im1;
    double2 d_im1 = make_double2((y_i.x - y_im1.x) * inv_him1, (y_i.y - y_im1.y) * inv_him1);
    double cp = 0.0;
    double2 dp = make_double2(0.0, 0.0);
    for (int64_t i = rs; i <= re; ++i) {
        const double x_ip1 = kn.x(i + 1);
        const double2 y_ip1 = Fg[(i + 1) * G];
        const double h_i = x_ip1 - x_i;
        const double inv_hi = 1.0 / h_i;
        const double2 d_i = make_double2((y_ip1.x - y_i.x) * inv_hi, (y_ip1.y - y_i.y) * inv_hi);
        double sub = h_im1, diag = 2.0 * (h_im1 + h_i), sup = h_i;
        if (i == 1 && true_lo) {          // not-a-knot at the left end, M_0 eliminated
            sub = 0.0;
            diag = (h_im1 + h_i) * (h_im1 + 2.0 * h_i) * inv_hi;
            sup = (h_i * h_i - h_im1 * h_im1) * inv_hi;
        }
        if (i == N - 2 && true_hi) {      // not-a-knot at the right end, M_{N-1} eliminated
            // (for N == 4 with i == 1 == N-2... cannot happen: N-2 == 2 there)
            const double dg = (h_im1 + h_i) * (2.0 * h_im1 + h_i) * inv_him1;
            const double sb = (h_im1 * h_im1 - h_i * h_i) * inv_him1;
            if (i == 1 && true_lo) {
                // N == 3 is rejected on the host, so this only guards the impossible
            }
            diag = dg;
            sub = sb;
            sup = 0.0;
        }
        if (i == rs && !true_lo) sub = 0.0;   // natural cut: M_lo = 0
        if (i == re && !true_hi) sup = 0.0;   // natural cut: M_hi = 0
        const double2 rhs = make_double2(6.0 * (d_i.x - d_im1.x), 6.0 * (d_i.y - d_im1.y));
        const double inv_den = 1.0 / (diag - sub * cp);
        cp = sup * inv_den;
        dp.x = (rhs.x - sub * dp.x) * inv_den;
        dp.y = (rhs.y - sub * dp.y) * inv_den;
        const int64_t s = i - lo;
        wc[s * G] = cp;
        wd[s * G] = dp;
        x_im1 = x_i; x_i = x_ip1;
        y_im1 = y_i; y_i = y_ip1;
        h_im1 = h_i; inv_him1 = inv_hi;
        d_im1 = d_i;
    }

    // ---------------- moment at the upper end of the window
    // after the loop: x_i = x(hi), y_i = F[hi], h_im1 = h_{hi-1}
    double2 M_ip1;   // M at knot i+1 during the backward sweep; starts as M_hi
    if (true_hi) {
        // M_{N-1} = ((h_{N-3}+h_{N-2}) M_{N-2} - h_{N-2} M_{N-3}) / h_{N-3}
        const double2 M_re = wd[(re - lo) * G];
        double2 M_rem1;
        {
            const double c1 = wc[(re - 1 - lo) * G];
            const double2 d1 = wd[(re - 1 - lo) * G];
            M_rem1 = make_double2(d1.x - c1 * M_re.x, d1.y - c1 * M_re.y);
            if (re - 1 < rs) {   // N == 4 with the window starting at 0: M_{N-3} = M_1 = M_re? no: re-1 = 1 = rs. never < rs.
                M_rem1 = M_re;
            }
        }
        const double hL = h_im1;                 // h_{N-2}
        const double hL1 = kn.x(N - 2) - kn.x(N - 3);   // h_{N-3}
        M_ip1.x = ((hL1 + hL) * M_re.x - hL * M_rem1.x) / hL1;
        M_ip1.y = ((hL1 + hL) * M_re.y - hL * M_rem1.y) / hL1;
    } else {
        M_ip1 = make_double2(0.0, 0.0);
    }

    // ---------------- output pointer: last output handled by this chunk
    int64_t ip = -1;
    if (MODE != 0) {
    } else if (last_chunk) {
        ip = Nout - 1;
    } else {
        const double xb = kn.x(b);
        int64_t lo_s = 0, hi_s = Nout;     // count of up[] < xb
        while (lo_s < hi_s) {
            int64_t mid = (lo_s + hi_s) >> 1;
            if (up[mid] < xb) lo_s = mid + 1; else hi_s = mid;
        }
        ip = lo_s - 1;
    }

    // ---------------- back substitution fused with evaluation
    double x_ip1 = x_i;          // x(hi)
    double2 y_ip1 = y_i;         // F[hi]
    double2 M_ip2 = make_double2(0.0, 0.0);
    for (int64_t i = re; i >= rs; --i) {
        const int64_t s = i - lo;
        const double c = wc[s * G];
        const double2 d = wd[s * G];
        double2 M_i;
        if (i == re) M_i = d;    // sup was zero or eliminated
        else M_i = make_double2(d.x - c * M_ip1.x, d.y - c * M_ip1.y);
        const double xi = kn.x(i);
        const double2 yi = Fg[i * G];
        if (i < b && i >= a) {
            const double h = x_ip1 - xi;
            const double inv_h = 1.0 / h;
            if (MODE == 0) {
                while (ip >= 0) {
                    const double u = up[ip];
                    if (u < xi && i > 0) break;
                    out[ip * G + g] = eval_piece(u, xi, x_ip1, h, inv_h, yi, y_ip1, M_i, M_ip1);
                    --ip;
                }
            } else if (MODE == 1) {
                const double h6 = h * (1.0 / 6.0);
                const double2 dl = make_double2((y_ip1.x - yi.x) * inv_h, (y_ip1.y - yi.y) * inv_h);
                out[i * G + g] = make_double2(dl.x - h6 * (2.0 * M_i.x + M_ip1.x), dl.y - h6 * (2.0 * M_i.y + M_ip1.y));
                if (i == N - 2)
                    out[(N - 1) * G + g] =
                        make_double2(dl.x + h6 * (M_i.x + 2.0 * M_ip1.x), dl.y + h6 * (M_i.y + 2.0 * M_ip1.y));
            } else {
                out[i * G + g] = M_i;
                if (i == N - 2) out[(N - 1) * G + g] = M_ip1;
            }
        }
        M_ip2 = M_ip1;
        M_ip1 = M_i;
        x_ip1 = xi;
        y_ip1 = yi;
    }
    // interval 0 (knots 0,1) when the window starts at the true left end
    if (true_lo && a == 0) {
        // here M_ip1 = M_1, M_ip2 = M_2, x_ip1 = x_1, y_ip1 = F[1]
        const double x0 = kn.x(0), x1 = x_ip1, x2 = kn.x(2);
        const double h0 = x1 - x0, h1 = x2 - x1;
        double2 M0;
        M0.x = ((h0 + h1) * M_ip1.x - h0 * M_ip2.x) / h1;
        M0.y = ((h0 + h1) * M_ip1.y - h0 * M_ip2.y) / h1;
        const double2 y0 = Fg[0];
        const double inv_h = 1.0 / h0;
        if (MODE == 0) {
            while (ip >= 0) {
                const double u = up[ip];
                out[ip * G + g] = eval_piece(u, x0, x1, h0, inv_h, y0, y_ip1, M0, M_ip1);
                --ip;
            }
        } else if (MODE == 1) {
            const double h6 = h0 * (1.0 / 6.0);
            const double2 dl = make_double2((y_ip1.x - y0.x) * inv_h, (y_ip1.y - y0.y) * inv_h);
            out[g] = make_double2(dl.x - h6 * (2.0 * M0.x + M_ip1.x), dl.y - h6 * (2.0 * M0.y + M_ip1.y));
        } else {
            out[g] = M0;
        }
    }
}

}  // namespace scrib200

extern "C" size_t scrib200_spline_remap_workspace_bytes(int64_t n_times, int G, int chunk) {
    using namespace scrib200;
    if (chunk <= 0) chunk = DEFAULT_CHUNK;
    if (n_times < 2) return 0;
    int64_t nchunks = (n_times - 1 + chunk - 1) / chunk;
    size_t W = (size_t)chunk + 2 * HALO + 2;
    return (size_t)nchunks * W * (size_t)G * 3 * sizeof(double) + 16;
}

extern "C" int scrib200_bms_spline_remap(const double* t, int64_t n_times, const double* F, int G, const double* kconf,
                                         const double* alpha, const double* uprm, int64_t n_out, double* out,
                                         int chunk, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(t && F && kconf && alpha && uprm && out && workspace, "bms_spline_remap: null pointer");
    SCRIB200_REQUIRE(n_times >= 4, "bms_spline_remap: a cubic interpolating spline needs at least 4 knots; got %lld",
                     (long long)n_times);
    SCRIB200_REQUIRE(G > 0, "bms_spline_remap: G=%d", G);
    SCRIB200_REQUIRE(aligned16(F) && aligned16(out) && aligned16(workspace), "bms_spline_remap: pointers must be 16-byte aligned");
    if (chunk <= 0) chunk = DEFAULT_CHUNK;
    const size_t need = scrib200_spline_remap_workspace_bytes(n_times, G, chunk);
    SCRIB200_REQUIRE(workspace_bytes >= need, "bms_spline_remap: workspace too small (%zu < %zu)", workspace_bytes, need);
    if (n_out <= 0) return SCRIB200_OK;
    const int64_t nchunks = (n_times - 1 + chunk - 1) / chunk;
    SCRIB200_REQUIRE(nchunks <= 65535, "bms_spline_remap: too many chunks (%lld); raise `chunk`", (long long)nchunks);
    const int W = chunk + 2 * HALO + 2;
    // workspace: [nchunks][W][G] doubles for c', then [nchunks][W][G] double2 for d'
    double* ws_c = reinterpret_cast<double*>(workspace);
    size_t nc = (size_t)nchunks * W * G;
    nc = (nc + 1) & ~(size_t)1;   // keep the double2 part 16-byte aligned
    double2* ws_d = reinterpret_cast<double2*>(ws_c + nc);
    dim3 grid((G + 63) / 64, (unsigned)nchunks);
    bms_spline_remap_kernel<0><<<grid, 64, 0, (cudaStream_t)stream>>>(
        t, n_times, reinterpret_cast<const double2*>(F), G, kconf, alpha, uprm, n_out, reinterpret_cast<double2*>(out),
        chunk, W, ws_c, ws_d);
    SCRIB200_CHECK_LAUNCH("bms_spline_remap");
    return SCRIB200_OK;
}

extern "C" int scrib200_spline_derivative(const double* t, int64_t n_times, const double* data, int ncol,
                                          const double* ones, const double* zeros, int order, double* out, int chunk,
                                          void* workspace, size_t workspace_bytes, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(t && data && ones && zeros && out && workspace, "spline_derivative: null pointer");
    SCRIB200_REQUIRE(n_times >= 4, "spline_derivative: a cubic spline needs at least 4 knots; got %lld",
                     (long long)n_times);
    SCRIB200_REQUIRE(order == 1 || order == 2, "spline_derivative: order must be 1 or 2");
    SCRIB200_REQUIRE(ncol > 0, "spline_derivative: ncol=%d", ncol);
    if (chunk <= 0) chunk = DEFAULT_CHUNK;
    const size_t need = scrib200_spline_remap_workspace_bytes(n_times, ncol, chunk);
    SCRIB200_REQUIRE(workspace_bytes >= need, "spline_derivative: workspace too small (%zu < %zu)", workspace_bytes, need);
    const int64_t nchunks = (n_times - 1 + chunk - 1) / chunk;
    SCRIB200_REQUIRE(nchunks <= 65535, "spline_derivative: too many chunks (%lld); raise `chunk`", (long long)nchunks);
    const int W = chunk + 2 * HALO + 2;
    double* ws_c = reinterpret_cast<double*>(workspace);
    size_t nc = (size_t)nchunks * W * ncol;
    nc = (nc + 1) & ~(size_t)1;
    double2* ws_d = reinterpret_cast<double2*>(ws_c + nc);
    dim3 grid((ncol + 63) / 64, (unsigned)nchunks);
    if (order == 1)
        bms_spline_remap_kernel<1><<<grid, 64, 0, (cudaStream_t)stream>>>(
            t, n_times, reinterpret_cast<const double2*>(data), ncol, ones, zeros, nullptr, 0,
            reinterpret_cast<double2*>(out), chunk, W, ws_c, ws_d);
    else
        bms_spline_remap_kernel<2><<<grid, 64, 0, (cudaStream_t)stream>>>(
            t, n_times, reinterpret_cast<const double2*>(data), ncol, ones, zeros, nullptr, 0,
            reinterpret_cast<double2*>(out), chunk, W, ws_c, ws_d);
    SCRIB200_CHECK_LAUNCH("spline_derivative");
    return SCRIB200_OK;
}
