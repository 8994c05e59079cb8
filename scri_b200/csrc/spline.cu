// K2 / K8: batched not-a-knot cubic splines along time - the BMS retarded-time remap and the time calculus.
//
// Replaces scri/waveform_grid.py:576-588: for every grid point g the reference builds two scipy
// InterpolatedUnivariateSpline objects (k=3 interpolating spline, FITPACK not-a-knot) on the knots
//     x_i = k[g] * (t[i] - alpha[g])
// for Re and Im of F[:, g] and evaluates them at the common output times u'; and
// scri/waveform_base.py:689-695 (CubicSpline(t, data).derivative(k)(t)) with k = 1, alpha = 0.
//
// The interpolating cubic spline is unique, so it is computed here in the moment (second-derivative)
// form.  Row i of the tridiagonal system, multiplied through by h_{i-1} h_i so that no division is
// needed to set it up:
//     h_{i-1}^2 h_i M_{i-1} + 2 (h_{i-1}+h_i) h_{i-1} h_i M_i + h_{i-1} h_i^2 M_{i+1}
//         = 6 (h_{i-1} (y_{i+1}-y_i) - h_i (y_i-y_{i-1}))
// with the not-a-knot conditions eliminated into the first and last rows, solved by a Thomas sweep
// (one reciprocal per row).  The knots are the reference's own rounded abscissae (x_i computed as
// mul(k, sub(t, alpha)) with no FMA contraction), because at t ~ 1e4 one ulp of x already moves the
// interpolant at the 1e-13 level.
//
// Parallelisation: one thread per (grid point, chunk of knots); lanes run along g so every access to
// the [time, g] arrays is coalesced.  A chunk solves its knots plus a halo of HALO knots on each side
// with a natural cut (M = 0); the tridiagonal inverse decays at least as fast as 2^-k (0.268^k on
// uniform knots), so HALO = 64 puts the cut below 1e-19 relative; chunks that reach a true end of the
// series apply the exact not-a-knot rows there.
//
// Memory: a plain two-pass Thomas solve would round-trip 24 bytes per (knot, g) of forward-sweep
// coefficients through HBM (measured: 10.3 GB of DRAM traffic for 2.0 GB of algorithmic bytes).  Here the
// sweep is CHECKPOINTED: the forward pass keeps only the state entering every block of CKB rows
// (24/CKB bytes per knot); the backward pass reloads a checkpoint, re-eliminates that block into
// registers, back-substitutes it and evaluates the outputs that fall into its intervals on the spot, so
// the moments are never stored either.  F is read twice (second read in the same order, block by
// block), the output written once.
#include <math_constants.h>

#include "common.cuh"

namespace scrib200 {

constexpr int HALO = 64;
constexpr int DEFAULT_CHUNK = 512;
constexpr int CKB = 4;   // rows per checkpoint block
constexpr int PFD = 1;          // blocks of rows in flight beyond the ones the current block needs
constexpr int RING = 16;        // rows per thread in the shared-memory ring: (PFD + 2) blocks of CKB rows + 2 start rows
constexpr int SPLINE_THREADS = 64;
static_assert((PFD + 2) * CKB + 2 <= RING, "ring too small");

// Each thread streams its own column of F through a private ring in shared memory with cp.async: rows are in
// flight PFD blocks ahead without costing registers, and only cp.async.wait_group (no CTA barrier) orders them.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

struct Knots {
    const double* t;
    double k, alpha;
    __device__ __forceinline__ double x(int i) const { return __dmul_rn(k, __dsub_rn(t[i], alpha)); }
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

// Rolling state of the forward elimination at row i (before the row is processed).
struct Sweep {
    double x_i, h_im1, cp;
    double2 y_i, dy_im1, dp;     // dy_im1 = y_i - y_{i-1}
    int N, rs, re;
    bool true_lo, true_hi;

    __device__ __forceinline__ void start(const Knots& kn, double2 y_im1, double2 y_i0, int i0, double cp0, double2 dp0) {
        const double x_im1 = kn.x(i0 - 1);
        x_i = kn.x(i0);
        y_i = y_i0;
        h_im1 = x_i - x_im1;
        dy_im1 = make_double2(y_i.x - y_im1.x, y_i.y - y_im1.y);
        cp = cp0;
        dp = dp0;
    }

    // eliminate row i given knot/value i+1; afterwards the state refers to row i+1.
    // PLAIN: the row is neither the first nor the last of the window (no end conditions to apply).
    template <bool PLAIN>
    __device__ __forceinline__ void step(int i, double x_ip1, double2 y_ip1) {
        const double h_i = x_ip1 - x_i;
        const double2 dy_i = make_double2(y_ip1.x - y_i.x, y_ip1.y - y_i.y);
        const double hh = h_im1 * h_i;
        double sub = h_im1 * hh, diag = 2.0 * (h_im1 + h_i) * hh, sup = hh * h_i;
        if (!PLAIN) {
            if (i == 1 && true_lo) {          // not-a-knot at the left end, M_0 eliminated
                sub = 0.0;
                diag = (h_im1 + h_i) * (h_im1 + 2.0 * h_i) * h_im1;
                sup = (h_i * h_i - h_im1 * h_im1) * h_im1;
            }
            if (i == N - 2 && true_hi) {      // not-a-knot at the right end, M_{N-1} eliminated
                diag = (h_im1 + h_i) * (2.0 * h_im1 + h_i) * h_i;
                sub = (h_im1 * h_im1 - h_i * h_i) * h_i;
                sup = 0.0;
            }
            if (i == rs && !true_lo) sub = 0.0;   // natural cut: M_lo = 0
            if (i == re && !true_hi) sup = 0.0;   // natural cut: M_hi = 0
        }
        const double2 rhs = make_double2(6.0 * (h_im1 * dy_i.x - h_i * dy_im1.x), 6.0 * (h_im1 * dy_i.y - h_i * dy_im1.y));
        const double inv_den = 1.0 / (diag - sub * cp);
        cp = sup * inv_den;
        dp.x = (rhs.x - sub * dp.x) * inv_den;
        dp.y = (rhs.y - sub * dp.y) * inv_den;
        x_i = x_ip1;
        y_i = y_ip1;
        h_im1 = h_i;
        dy_im1 = dy_i;
    }
};

// MODE 0: evaluate at up[] (the BMS remap); MODE 1 / 2: first / second derivative at the knots themselves
// (out is then [N, G]; `up`/`Nout` unused).
template <int MODE>
__global__ void __launch_bounds__(64, 8)
spline_ckpt_kernel(const double* __restrict__ t, int N, const double2* __restrict__ F, int G,
                   const double* __restrict__ kconf, const double* __restrict__ alpha, const double* __restrict__ up,
                   int Nout, double2* __restrict__ out, int tshift, int C, int NCK, double* __restrict__ ws_c,
                   double2* __restrict__ ws_d) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;   // threads with g >= G only help load s_t[], then leave
    // MODE 0 output layout: tshift == 0 -> time-major out[i'*G + g]; tshift = log2(T) > 0 -> time-tiled
    // out[(i'/T)*(G*T) + g*T + i'%T].  Tiling makes each thread's consecutive outputs contiguous (runs of T), so L2
    // merges them into full sectors (time-major 16-byte stores from lanes sitting at different i' never meet their
    // sector neighbour: 2x write traffic), and hands the analysis kernel one contiguous [G, T] tile per CTA.
    const int tmask = (1 << tshift) - 1;
    const int64_t tileGT = (int64_t)G << tshift;
    double2* og = out + (tshift > 0 ? ((int64_t)g << tshift) : (int64_t)g);
#define SCRIB200_OIDX(ip_) (tshift > 0 ? (int64_t)((ip_) >> tshift) * tileGT + ((ip_) & tmask) : (int64_t)(ip_) * G)
    const int chunk = blockIdx.y;
    const int a = chunk * C;                                 // first interval of this chunk
    const int b = (a + C < N - 1) ? a + C : N - 1;           // intervals a .. b-1, knots a .. b
    const bool last_chunk = (b == N - 1);
    const int lo = (a - HALO > 0) ? a - HALO : 0;
    const int hi = (b + HALO < N - 1) ? b + HALO : N - 1;
    const bool true_lo = (lo == 0), true_hi = (hi == N - 1);
    const int rs = lo + 1, re = hi - 1;                      // rows solved (N >= 4 => re >= rs + 1)
    const int nrows = re - rs + 1;
    const int nblocks = (nrows + CKB - 1) / CKB;
    const int first = nrows - (nblocks - 1) * CKB;           // size of block 0 (1..CKB); the others are full
    // first block the backward pass has to visit: the one holding row max(a, rs)
    int kmin = 0;
    if (a > rs) {
        const int j = a - rs;
        kmin = (j < first) ? 0 : 1 + (j - first) / CKB;
    }

    // the window's sample times go to shared memory once: every x_i below is then an LDS, not an L2 round trip
    extern __shared__ __align__(16) double2 s_ring[];                  // [RING][SPLINE_THREADS]
    double* s_t = reinterpret_cast<double*>(s_ring + RING * SPLINE_THREADS);
    for (int j = threadIdx.x; j <= hi - lo; j += blockDim.x) s_t[j] = t[lo + j];
    __syncthreads();
    if (g >= G) return;
    Knots kn{s_t - lo, kconf[g], alpha[g]};
    double2* my_ring = s_ring + threadIdx.x;
#define RING_AT(row_) my_ring[((row_) & (RING - 1)) * SPLINE_THREADS]
    double* wc = ws_c + ((size_t)chunk * NCK) * G + g;       // checkpoint k -> wc[k*G]
    double2* wd = ws_d + ((size_t)chunk * NCK) * G + g;
    const double2* Fg = F + g;

    Sweep sw;
    sw.N = N; sw.rs = rs; sw.re = re; sw.true_lo = true_lo; sw.true_hi = true_hi;

    // block k covers rows blk_i0(k) .. blk_i0(k)+blk_nb(k)-1 and needs F rows up to blk_i0(k)+blk_nb(k)
#define BLK_I0(k_) ((k_) == 0 ? rs : rs + first + ((k_) - 1) * CKB)
#define BLK_NB(k_) ((k_) == 0 ? first : CKB)
    // issue the cp.async copies of the F rows block k_ brings in: rows i0+1 .. i0+nb (plus lo, rs for block 0)
    auto issue_block = [&](int kk) {
        if (kk >= 0 && kk < nblocks) {
            const int i0 = BLK_I0(kk), nb = BLK_NB(kk);
            if (kk == 0) {
                cp_async16(&RING_AT(lo), Fg + (int64_t)lo * G);
                cp_async16(&RING_AT(rs), Fg + (int64_t)rs * G);
            }
#pragma unroll
            for (int r = 0; r < CKB; ++r)
                if (r < nb) cp_async16(&RING_AT(i0 + r + 1), Fg + (int64_t)(i0 + r + 1) * G);
        }
        cp_async_commit();   // always commit, so the group count stays in step with the block count
    };

    // ---------------- forward pass: only the checkpoints survive
#pragma unroll
    for (int kk = 0; kk <= PFD; ++kk) issue_block(kk < nblocks - 1 ? kk : -1);
    for (int k = 0; k + 1 < nblocks; ++k) {
        const int i0 = BLK_I0(k);
        const int nb = BLK_NB(k);
        cp_async_wait<PFD>();                // everything up to block k has landed (PFD younger groups may be in flight)
        if (k == 0) sw.start(kn, RING_AT(lo), RING_AT(rs), rs, 0.0, make_double2(0.0, 0.0));
        if (k > 0) {   // full block strictly inside the window (the last block is never processed here)
#pragma unroll
            for (int r = 0; r < CKB; ++r) sw.step<true>(i0 + r, kn.x(i0 + r + 1), RING_AT(i0 + r + 1));
        } else {
#pragma unroll
            for (int r = 0; r < CKB; ++r)
                if (r < nb) sw.step<false>(i0 + r, kn.x(i0 + r + 1), RING_AT(i0 + r + 1));
        }
        if (k + 1 >= kmin) {
            wc[(size_t)(k + 1) * G] = sw.cp;
            wd[(size_t)(k + 1) * G] = sw.dp;
        }
        // the slots of rows <= i0+nb are free now (row i0+nb lives on in sw.y_i): bring in block k+PFD+1
        issue_block(k + PFD + 1 < nblocks - 1 ? k + PFD + 1 : -1);
    }
    cp_async_wait<0>();

    // ---------------- output pointer: last output handled by this chunk
    int ip = -1;
    if (MODE != 0) {
    } else if (last_chunk) {
        ip = Nout - 1;
    } else {
        const double xb_ = kn.x(b);
        int lo_s = 0, hi_s = Nout;     // count of up[] < x_b
        while (lo_s < hi_s) {
            int mid = (lo_s + hi_s) >> 1;
            if (up[mid] < xb_) lo_s = mid + 1; else hi_s = mid;
        }
        ip = lo_s - 1;
    }
    // u_cur = up[ip], u_nxt = up[ip-1]: fetched two outputs ahead so the interval test never waits on memory
    double u_cur = (MODE == 0 && ip >= 0) ? up[ip] : -CUDART_INF;
    double u_nxt = (MODE == 0 && ip >= 1) ? up[ip - 1] : -CUDART_INF;

    // ---------------- backward pass, block by block from the right
    double2 M_ip1 = make_double2(0.0, 0.0), M_ip2 = make_double2(0.0, 0.0);
    double x_ip1 = 0.0;
    double2 y_ip1 = make_double2(0.0, 0.0);
    // block k needs its own rows and (for the two rows below i0) block k-1's: PFD + 2 groups go out first
#pragma unroll
    for (int d = 0; d <= PFD + 1; ++d) issue_block(nblocks - 1 - d >= kmin - 1 ? nblocks - 1 - d : -1);
    double cpn = 0.0;                              // checkpoint of the block being processed, fetched one block early
    double2 dpn = make_double2(0.0, 0.0);
    if (nblocks - 1 > 0) {
        cpn = wc[(size_t)(nblocks - 1) * G];
        dpn = wd[(size_t)(nblocks - 1) * G];
    }
    for (int k = nblocks - 1; k >= kmin; --k) {
        const int i0 = BLK_I0(k);
        const int nb = BLK_NB(k);
        const double cp0 = cpn;
        const double2 dp0 = dpn;
        if (k - 1 >= kmin && k - 1 > 0) {
            cpn = wc[(size_t)(k - 1) * G];
            dpn = wd[(size_t)(k - 1) * G];
        } else {
            cpn = 0.0;
            dpn = make_double2(0.0, 0.0);
        }
        cp_async_wait<PFD>();                      // blocks k and k-1 have landed
        double cpb[CKB];
        double2 dpb[CKB];
#define XB(r_) kn.x(i0 + (r_) + 1)
#define YB(r_) RING_AT(i0 + (r_) + 1)
        sw.start(kn, RING_AT(i0 - 1), RING_AT(i0), i0, cp0, dp0);
        const double x_i0 = sw.x_i;
        const double2 y_i0 = sw.y_i;
        if (k > 0 && k < nblocks - 1) {
#pragma unroll
            for (int r = 0; r < CKB; ++r) {
                sw.step<true>(i0 + r, XB(r), YB(r));
                cpb[r] = sw.cp;
                dpb[r] = sw.dp;
            }
        } else {
#pragma unroll
            for (int r = 0; r < CKB; ++r)
                if (r < nb) {
                    sw.step<false>(i0 + r, XB(r), YB(r));
                    cpb[r] = sw.cp;
                    dpb[r] = sw.dp;
                }
        }
        if (k == nblocks - 1) {
            // moment at the upper end of the window (knot hi = re + 1); the last block has >= 2 rows
            if (true_hi) {
                double2 M_re = make_double2(0.0, 0.0), M_rem1 = make_double2(0.0, 0.0);
#pragma unroll
                for (int r = 0; r < CKB; ++r)
                    if (r == nb - 1) M_re = dpb[r];
#pragma unroll
                for (int r = 0; r < CKB; ++r)
                    if (r == nb - 2) M_rem1 = make_double2(dpb[r].x - cpb[r] * M_re.x, dpb[r].y - cpb[r] * M_re.y);
                const double hL = sw.h_im1;                        // h_{N-2}
                const double hL1 = kn.x(N - 2) - kn.x(N - 3);      // h_{N-3}
                M_ip1.x = ((hL1 + hL) * M_re.x - hL * M_rem1.x) / hL1;
                M_ip1.y = ((hL1 + hL) * M_re.y - hL * M_rem1.y) / hL1;
            }
            x_ip1 = sw.x_i;   // x(hi)
            y_ip1 = sw.y_i;   // F[hi]
        }
#pragma unroll
        for (int r = CKB - 1; r >= 0; --r) {
            if (r < nb) {
                const int i = i0 + r;
                double2 M_i;
                if (i == re) M_i = dpb[r];
                else M_i = make_double2(dpb[r].x - cpb[r] * M_ip1.x, dpb[r].y - cpb[r] * M_ip1.y);
                const double xi = (r == 0) ? x_i0 : XB(r - 1);
                const double2 yi = (r == 0) ? y_i0 : YB(r - 1);
                if (i < b && i >= a) {
                    const double h = x_ip1 - xi;
                    const double inv_h = 1.0 / h;
                    if (MODE == 0) {
                        while (u_cur >= xi) {
                            og[SCRIB200_OIDX(ip)] = eval_piece(u_cur, xi, x_ip1, h, inv_h, yi, y_ip1, M_i, M_ip1);
                            --ip;
                            u_cur = u_nxt;
                            u_nxt = (ip >= 1) ? up[ip - 1] : -CUDART_INF;
                        }
                    } else if (MODE == 1) {
                        const double h6 = h * (1.0 / 6.0);
                        const double2 dl = make_double2((y_ip1.x - yi.x) * inv_h, (y_ip1.y - yi.y) * inv_h);
                        out[(int64_t)i * G + g] = make_double2(dl.x - h6 * (2.0 * M_i.x + M_ip1.x), dl.y - h6 * (2.0 * M_i.y + M_ip1.y));
                        if (i == N - 2)
                            out[(int64_t)(N - 1) * G + g] =
                                make_double2(dl.x + h6 * (M_i.x + 2.0 * M_ip1.x), dl.y + h6 * (M_i.y + 2.0 * M_ip1.y));
                    } else {
                        out[(int64_t)i * G + g] = M_i;
                        if (i == N - 2) out[(int64_t)(N - 1) * G + g] = M_ip1;
                    }
                }
                M_ip2 = M_ip1;
                M_ip1 = M_i;
                x_ip1 = xi;
                y_ip1 = yi;
            }
        }
        // rows above i0 are dead now (block k-1 needs nothing above its own i0+nb = this i0): refill their slots
        issue_block(k - 2 - PFD >= kmin - 1 ? k - 2 - PFD : -1);
    }
    cp_async_wait<0>();
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
                og[SCRIB200_OIDX(ip)] = eval_piece(u_cur, x0, x1, h0, inv_h, y0, y_ip1, M0, M_ip1);
                --ip;
                u_cur = u_nxt;
                u_nxt = (ip >= 1) ? up[ip - 1] : -CUDART_INF;
            }
        } else if (MODE == 1) {
            const double h6 = h0 * (1.0 / 6.0);
            const double2 dl = make_double2((y_ip1.x - y0.x) * inv_h, (y_ip1.y - y0.y) * inv_h);
            out[g] = make_double2(dl.x - h6 * (2.0 * M0.x + M_ip1.x), dl.y - h6 * (2.0 * M0.y + M_ip1.y));
        } else {
            out[g] = M0;
        }
    }
#undef SCRIB200_OIDX
}

static inline int checkpoints_per_chunk(int chunk) { return (chunk + 2 * HALO + 2 + CKB - 1) / CKB + 2; }

template <int MODE>
static int launch_spline(const double* t, int64_t n_times, const double* F, int G, const double* kconf,
                         const double* alpha, const double* uprm, int64_t n_out, double* out, int tshift, int chunk,
                         void* workspace, size_t workspace_bytes, void* stream, const char* name) {
    if (chunk <= 0) chunk = DEFAULT_CHUNK;
    SCRIB200_REQUIRE(n_times < (int64_t)2147483000 && n_out < (int64_t)2147483000, "%s: series longer than 2^31 samples", name);
    const size_t need = scrib200_spline_remap_workspace_bytes(n_times, G, chunk);
    SCRIB200_REQUIRE(workspace_bytes >= need, "%s: workspace too small (%zu < %zu)", name, workspace_bytes, need);
    const int64_t nchunks = (n_times - 1 + chunk - 1) / chunk;
    SCRIB200_REQUIRE(nchunks <= 65535, "%s: too many chunks (%lld); raise `chunk`", name, (long long)nchunks);
    const int NCK = checkpoints_per_chunk(chunk);
    // workspace: [nchunks][NCK][G] doubles for c', then [nchunks][NCK][G] double2 for d'
    double* ws_c = reinterpret_cast<double*>(workspace);
    size_t nc = (size_t)nchunks * NCK * G;
    nc = (nc + 1) & ~(size_t)1;   // keep the double2 part 16-byte aligned
    double2* ws_d = reinterpret_cast<double2*>(ws_c + nc);
    dim3 grid((G + 63) / 64, (unsigned)nchunks);
    // per-thread row ring + the window's sample times
    const size_t smem = (size_t)RING * SPLINE_THREADS * sizeof(double2) + (size_t)(chunk + 2 * HALO + 2) * sizeof(double);
    SCRIB200_REQUIRE(smem <= 200 * 1024, "%s: chunk=%d too large for the shared-memory time window", name, chunk);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(spline_ckpt_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    spline_ckpt_kernel<MODE><<<grid, 64, smem, (cudaStream_t)stream>>>(
        t, (int)n_times, reinterpret_cast<const double2*>(F), G, kconf, alpha, uprm, (int)n_out,
        reinterpret_cast<double2*>(out), tshift, chunk, NCK, ws_c, ws_d);
    SCRIB200_CHECK_LAUNCH(name);
    return SCRIB200_OK;
}

}  // namespace scrib200

extern "C" size_t scrib200_spline_remap_workspace_bytes(int64_t n_times, int G, int chunk) {
    using namespace scrib200;
    if (chunk <= 0) chunk = DEFAULT_CHUNK;
    if (n_times < 2) return 16;
    int64_t nchunks = (n_times - 1 + chunk - 1) / chunk;
    return (size_t)nchunks * checkpoints_per_chunk(chunk) * (size_t)G * 3 * sizeof(double) + 16;
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
    if (n_out <= 0) return SCRIB200_OK;
    return launch_spline<0>(t, n_times, F, G, kconf, alpha, uprm, n_out, out, 0, chunk, workspace, workspace_bytes, stream,
                            "bms_spline_remap");
}

extern "C" int scrib200_bms_spline_remap_tiled(const double* t, int64_t n_times, const double* F, int G,
                                               const double* kconf, const double* alpha, const double* uprm,
                                               int64_t n_out, double* out, int tile, int chunk, void* workspace,
                                               size_t workspace_bytes, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(t && F && kconf && alpha && uprm && out && workspace, "bms_spline_remap_tiled: null pointer");
    SCRIB200_REQUIRE(n_times >= 4, "bms_spline_remap_tiled: a cubic interpolating spline needs at least 4 knots; got %lld",
                     (long long)n_times);
    int tshift = 0;
    while ((1 << tshift) < tile) ++tshift;
    SCRIB200_REQUIRE(G > 0 && tile >= 2 && (1 << tshift) == tile, "bms_spline_remap_tiled: G=%d, tile=%d must be a power of two >= 2",
                     G, tile);
    SCRIB200_REQUIRE(aligned16(F) && aligned16(out) && aligned16(workspace),
                     "bms_spline_remap_tiled: pointers must be 16-byte aligned");
    if (n_out <= 0) return SCRIB200_OK;
    return launch_spline<0>(t, n_times, F, G, kconf, alpha, uprm, n_out, out, tshift, chunk, workspace, workspace_bytes,
                            stream, "bms_spline_remap_tiled");
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
    SCRIB200_REQUIRE(aligned16(data) && aligned16(out) && aligned16(workspace), "spline_derivative: pointers must be 16-byte aligned");
    if (order == 1)
        return launch_spline<1>(t, n_times, data, ncol, ones, zeros, nullptr, 0, out, 0, chunk, workspace, workspace_bytes,
                                stream, "spline_derivative");
    return launch_spline<2>(t, n_times, data, ncol, ones, zeros, nullptr, 0, out, 0, chunk, workspace, workspace_bytes,
                            stream, "spline_derivative");
}
