// K2 / K8 (second generation): batched not-a-knot cubic splines along time, tile-resident in shared memory.
//
// Replaces scri/waveform_grid.py:576-588 (two scipy InterpolatedUnivariateSpline objects per grid point on the knots
// x_i = k[g] (t[i] - alpha[g]), evaluated at the common output times u') and scri/waveform_base.py:689-703
// (CubicSpline(t, data).derivative(k)(t), .antiderivative(k)(t)).
//
// Three observations shape the kernel:
//  1. The knots of every grid point are an affine image of the same sample times t, and the interpolating cubic
//     spline is affine-invariant.  In t-units the moment system  T(t) M = 6 [y_{i-1}, y_i, y_{i+1}]  has ONE matrix
//     for all columns, so its Thomas factorisation is done once (`spline_factor_kernel`, one thread per row; the
//     c' recurrence contracts by <= 1/12 per row, so 40 rows of run-in reproduce it to rounding) and stored as a
//     table per row.  The per-column work is then division free:
//         d_i = P_i (y_{i+1}-y_i) - Q_i (y_i-y_{i-1}) - W_i d_{i-1}        (forward)
//         M_i = d_i - c'_i M_{i+1}                                         (backward)
//     The evaluation still uses the reference's own rounded abscissae x_i = fl(k fl(t_i - alpha)) for the position
//     inside the interval (one ulp of x at t ~ 1e4 moves the interpolant at the 1e-13 level); only the small
//     curvature term uses the shared factorisation (h_t^2 M_t = h_x^2 M_x up to 1e-11 relative, i.e. < 1e-14 of the
//     value).
//  2. Both recurrences are linear, so rows are cut into global blocks of 16: a half-warp (16 real columns) sweeps one
//     block from a zero start, and the exact coupling to the neighbouring blocks is restored with the column-independent
//     products Phi_i = prod(-W) (from the block start to i) and Psi_i = prod(-c') (from i to the block end), which
//     are table entries too:  d_i = d0_i + Phi_i d_{start-1},  M_i = m0_i + Psi_i M_{end+1}.  All blocks of a tile run
//     in parallel with no run-in.
//  3. Both recurrences forget their start geometrically (|W| ~ 0.27 on uniform samples), so a tile of `body`
//     intervals plus `halo` rows on each side is self-contained (the halo is chosen from the measured decay).  A CTA
//     owns [body + 2 halo + 2] rows x 16 real columns (8 complex grid points = 128-byte rows) of F in shared memory,
//     read once, and evaluates the outputs that fall into its intervals with lanes running along the output index,
//     so the stores into the time-tiled output are full 128-byte runs.  No workspace, F is never re-read from HBM
//     (the halo rows of neighbouring tiles come from L2).
#include <cuda.h>
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"

namespace scrib200 {

constexpr int ST_COLS = 16;            // real columns per CTA (8 complex grid points: one 128-byte row segment)
constexpr int ST_ROWD = 16;            // doubles per shared-memory row: the TMA box row, stored with the 128-byte swizzle
constexpr int ST_BR = 16;              // rows per sweep block (global alignment)
constexpr int ST_TAB6 = 6;             // doubles per row in use: P, Q, Phi, c', W, Psi  (tab[i * 6 + f], rows contiguous)
constexpr int FACTOR_RUNIN = 40;
constexpr int ST_MAXBLK = 24;          // max sweep blocks per tile (384 threads)

__device__ __forceinline__ unsigned st_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// Shared-memory position (in doubles) of real column c of tile row `rrel` under the TMA 128-byte swizzle: the 16-byte
// chunk index of a row is XORed with the row index modulo 8, so the 16-byte gathers of the evaluation (one chunk, many
// rows) and the row-wise sweeps (one row, all chunks) are both bank-conflict free on a dense 128-byte pitch.
__device__ __forceinline__ int st_idx(int rrel, int c) { return rrel * ST_ROWD + ((((c >> 1) ^ (rrel & 7)) << 1) | (c & 1)); }

// Row r (1 <= r <= N-2) of the moment system in t-units, not-a-knot conditions eliminated into rows 1 and N-2.
__device__ __forceinline__ void nak_row(const double* __restrict__ t, int N, int r, double& sub, double& diag, double& sup,
                                        double& hm, double& hp) {
    hm = t[r] - t[r - 1];
    hp = t[r + 1] - t[r];
    sub = hm;
    diag = 2.0 * (hm + hp);
    sup = hp;
    if (r == 1) {   // M_0 = ((h0+h1) M_1 - h0 M_2) / h1
        sub = 0.0;
        diag = (hm + hp) * (hm + 2.0 * hp) / hp;
        sup = (hp * hp - hm * hm) / hp;
    }
    if (r == N - 2) {   // M_{N-1} = ((hm+hp) M_{N-2} - hp M_{N-3}) / hm      (N >= 4, so this is never row 1)
        diag = (hm + hp) * (2.0 * hm + hp) / hm;
        sub = (hm * hm - hp * hp) / hm;
        sup = 0.0;
    }
}

// tab[i] = (P, Q, Phi, c', W, Psi) for rows 1..N-2 (rows 0 and N-1 zeros), six doubles per row, rows contiguous; also
// u'_i = inv_gamma (t_i - tt) and the retained output block [lo, hi) = { i : umin <= u'_i <= umax }
// (waveform_grid.py:564-568), umin/umax reduced here from k (t_0 - alpha), k (t_{N-1} - alpha) over the G grid points.
__global__ void __launch_bounds__(256)
spline_factor_kernel(const double* __restrict__ t, int N, double* __restrict__ tab, double inv_gamma, int divide, double tt,
                     const double* __restrict__ kconf, const double* __restrict__ alpha, int G,
                     double* __restrict__ uprm, double* __restrict__ info) {
    __shared__ double s_red[2][8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (uprm != nullptr) {
        // every CTA reduces the (tiny) grid tables itself: no second launch, no host round trip
        const double t0 = t[0], t1 = t[N - 1];
        double umin = -CUDART_INF, umax = CUDART_INF;
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            const double k = kconf[g], al = alpha[g];
            umin = fmax(umin, __dmul_rn(k, __dsub_rn(t0, al)));
            umax = fmin(umax, __dmul_rn(k, __dsub_rn(t1, al)));
        }
        for (int o = 16; o > 0; o >>= 1) {
            umin = fmax(umin, __shfl_xor_sync(0xffffffffu, umin, o));
            umax = fmin(umax, __shfl_xor_sync(0xffffffffu, umax, o));
        }
        if ((threadIdx.x & 31) == 0) {
            s_red[0][threadIdx.x >> 5] = umin;
            s_red[1][threadIdx.x >> 5] = umax;
        }
        __syncthreads();
        umin = s_red[0][0];
        umax = s_red[1][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            umin = fmax(umin, s_red[0][w]);
            umax = fmin(umax, s_red[1][w]);
        }
        if (i < N) {
            // waveform_grid.py:565 multiplies by 1/gamma, transformations.py:393 divides by gamma: both are reproduced
            const double u = divide ? __ddiv_rn(__dsub_rn(t[i], tt), inv_gamma) : __dmul_rn(inv_gamma, __dsub_rn(t[i], tt));
            uprm[i] = u;
            // u' is non-decreasing: exactly one i starts / ends the retained block
            double up_prev = -CUDART_INF;
            if (i > 0) up_prev = divide ? __ddiv_rn(__dsub_rn(t[i - 1], tt), inv_gamma) : __dmul_rn(inv_gamma, __dsub_rn(t[i - 1], tt));
            if (u >= umin && !(up_prev >= umin)) info[0] = (double)i;          // lo: first u' >= umin
            if (u > umax && !(up_prev > umax)) info[1] = (double)i;            // hi: first u' >  umax
            if (i == 0) {
                info[4] = umin;
                info[5] = umax;
            }
        }
    }
    {   // info[6] = smallest sample spacing (the host sizes the halo of its pipelines with it): positive doubles order like their bits
        double h = (i < N - 1) ? t[i + 1] - t[i] : CUDART_INF;
        for (int o = 16; o > 0; o >>= 1) h = fmin(h, __shfl_xor_sync(0xffffffffu, h, o));
        if ((threadIdx.x & 31) == 0 && h > 0.0 && h < CUDART_INF)
            atomicMin(reinterpret_cast<unsigned long long*>(info + 6), (unsigned long long)__double_as_longlong(h));
    }
    if (i >= N) return;
    double row[ST_TAB6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (i >= 1 && i <= N - 2) {
        const int bstart = (i & ~(ST_BR - 1)) > 1 ? (i & ~(ST_BR - 1)) : 1;
        const int bend = ((i | (ST_BR - 1)) < N - 2) ? (i | (ST_BR - 1)) : N - 2;
        int r = i - FACTOR_RUNIN;
        if (r < 1) r = 1;
        double cp = 0.0, phi = 1.0, psi = 1.0;
        for (; r <= bend; ++r) {
            double sub, diag, sup, hm, hp;
            nak_row(t, N, r, sub, diag, sup, hm, hp);
            const double e = 1.0 / (diag - sub * cp);
            cp = sup * e;
            const double W = sub * e;
            if (r >= bstart && r <= i) phi *= -W;
            if (r >= i) psi *= -cp;
            if (r == i) {
                row[0] = 6.0 * e / hp;
                row[1] = 6.0 * e / hm;
                row[4] = W;
                row[3] = cp;
            }
        }
        row[2] = phi;
        row[5] = psi;
    }
    double2* dst = reinterpret_cast<double2*>(tab + (size_t)i * ST_TAB6);
    dst[0] = make_double2(row[0], row[1]);
    dst[1] = make_double2(row[2], row[3]);
    dst[2] = make_double2(row[4], row[5]);
}

__global__ void spline_info_init_kernel(double* __restrict__ info, double n) {
    if (threadIdx.x < 8) info[threadIdx.x] = (threadIdx.x < 2) ? n : (threadIdx.x == 6 ? CUDART_INF : 0.0);
}

// Worst decay of the two recurrences over any window of 32 / 64 consecutive rows:
// info[2] = max_i prod_{r=i-31..i} |W_r| (forward) or prod |c'_r| (backward), info[3] the same for 64 rows.
__global__ void __launch_bounds__(256)
spline_decay_kernel(const double* __restrict__ tab, int N, double* __restrict__ info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double d32 = 0.0, d64 = 0.0;
    if (i >= 64 && i <= N - 2) {
        double pw = 1.0, pc = 1.0;
        for (int r = i; r > i - 64; --r) {
            pw *= fabs(tab[(size_t)r * ST_TAB6 + 4]);
            pc *= fabs(tab[(size_t)r * ST_TAB6 + 3]);
            if (r == i - 31) d32 = fmax(pw, pc);
        }
        d64 = fmax(pw, pc);
    }
    for (int o = 16; o > 0; o >>= 1) {
        d32 = fmax(d32, __shfl_xor_sync(0xffffffffu, d32, o));
        d64 = fmax(d64, __shfl_xor_sync(0xffffffffu, d64, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long*>(info + 2), (unsigned long long)__double_as_longlong(d32));
        atomicMax(reinterpret_cast<unsigned long long*>(info + 3), (unsigned long long)__double_as_longlong(d64));
    }
}

// J[tile][g] = first output j with up[j] >= x_{tile*body}(g) (J[0][g] = 0, J[ntiles][g] = Nout): the outputs column g's
// tile `tile` evaluates are J[tile][g] .. J[tile+1][g]-1.  One thread per entry: the 17 dependent probes of each binary
// search are hidden by the parallelism here instead of sitting on every tile CTA's critical path.
__global__ void __launch_bounds__(256)
spline_jrange_kernel(const double* __restrict__ t, int N, int G, const double* __restrict__ kconf,
                     const double* __restrict__ alpha, const double* __restrict__ up, int Nout, int body, int ntiles,
                     int* __restrict__ J) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (ntiles + 1) * G) return;
    const int tile = idx / G, g = idx - tile * G;
    int j = (tile == 0) ? 0 : Nout;
    if (tile > 0 && tile < ntiles) {
        const double xb = __dmul_rn(kconf[g], __dsub_rn(t[tile * body], alpha[g]));
        int lo_s = 0, hi_s = Nout;
        while (lo_s < hi_s) {
            const int mid = (lo_s + hi_s) >> 1;
            if (up[mid] < xb) lo_s = mid + 1; else hi_s = mid;
        }
        j = lo_s;
    }
    J[idx] = j;
}

// flags[tile][column group] = 1 when any of the group's grid points has an output in the tile.  A call that evaluates a
// slice of the output times (a slab of u', an interpolation onto a few points) then skips the tiles it does not touch
// instead of staging and sweeping them.
__global__ void __launch_bounds__(256)
spline_tile_flags_kernel(const int* __restrict__ J, int ntiles, int G, int ncg, int* __restrict__ flags) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ntiles * ncg) return;
    const int tile = idx / ncg, cg = idx - tile * ncg;
    int any = 0;
    for (int c = 0; c < ST_COLS / 2; ++c) {
        const int g = cg * (ST_COLS / 2) + c;
        if (g < G) any |= (J[(size_t)tile * G + g] < J[(size_t)(tile + 1) * G + g]);
    }
    flags[idx] = any;
}

// MODE 0: evaluate at up[] (the BMS remap); MODE 1 / 2: first / second derivative at the knots; MODE 3: the
// definite integral over [t_i, t_{i+1}] stored at row i+1 (row 0 = 0) - a column scan turns it into the antiderivative;
// MODE 4: the increment of the second antiderivative, `up` then holding the (scanned) first antiderivative [N, G].
// For MODE >= 1 `out` is [N, G] complex time-major and Nout/tshift are unused.
// blockDim.x >= 16 * (body + 2 halo) / 16 (one half-warp per sweep block), body and halo multiples of 16.
//
// Staging: the F rows of the tile arrive as two TMA boxes (`tmF`: F viewed as a [rows, 2G] FP64 tensor, box = 16 columns x
// box_rows rows, 128-byte swizzle, columns / rows beyond the tensor zero-filled by the hardware) and the table rows as one
// bulk copy, all completing on one mbarrier: no LSU instructions and no shared-memory store wavefronts are spent on
// staging (the cp.async version of this kernel spent 40 % of its shared-memory cycles and 25 % of its instructions there).
template <int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, 2)
spline_tile_kernel(const __grid_constant__ CUtensorMap tmF, const double* __restrict__ t, int N, int G,
                   const double* __restrict__ kconf, const double* __restrict__ alpha,
                   const double* __restrict__ tab, const double* __restrict__ up, int Nout,
                   double* __restrict__ out, int tshift, int body, int halo, const int* __restrict__ J,
                   const int* __restrict__ flags, int box_rows, int tile_off) {
    const int bx = blockIdx.x, by = blockIdx.y + tile_off, bz = blockIdx.z;
    if (MODE == 0 && flags[by * gridDim.x + bx] == 0) return;   // no output time falls in this tile
    extern __shared__ unsigned char s_raw[];
    __shared__ __align__(8) unsigned long long s_mbar;

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int G2 = 2 * G;
    const int col0 = bx * ST_COLS;                          // first real column of this CTA
    const int a = by * body;                                // intervals a .. b-1, knots a .. b
    const int b = (a + body < N - 1) ? a + body : N - 1;
    const bool last_tile = (b == N - 1);
    const int slo = (a - halo > 1) ? a - halo : 1;          // rows slo .. shi are swept
    const int shi = (b + halo - 1 < N - 2) ? b + halo - 1 : N - 2;
    const int ylo = slo - 1, yhi = shi + 1;                 // rows the tile needs
    const int ybase = ylo & ~7;                             // row held by shared-memory row 0: keeps (r - ybase) & 7 == r & 7
    const int nrows = 2 * box_rows;                         // shared-memory rows (>= yhi - ybase + 1)
    const int qlo = slo / ST_BR, qhi = shi / ST_BR;         // sweep blocks of this tile
    const int nblk_raw = (body + 2 * halo) / ST_BR;
    const int nblk_max = nblk_raw > ST_BR ? nblk_raw : ST_BR;   // the antiderivative modes keep ST_BR strip totals in sEdgeD

    double* sY = reinterpret_cast<double*>(s_raw + ((1024u - (st_smem_u32(s_raw) & 1023u)) & 1023u));   // [nrows][16] F (swizzled)
    double* sD = sY + (size_t)nrows * ST_ROWD;              // [nrows][16] moments M (same layout)
    double* sTab = sD + (size_t)nrows * ST_ROWD;            // [nrows][6]  P, Q, Phi, c', W, Psi     row r -> r - ybase
    double* sT = sTab + (size_t)nrows * ST_TAB6;            // [nrows]     sample times
    double* sEdgeD = sT + nrows;                            // [nblk_max][ST_COLS] d0 at the block ends
    double* sEdgeM = sEdgeD + nblk_max * ST_COLS;           // [nblk_max][ST_COLS] m0 at the block starts
    // per-column constants of the evaluation, addressed from one register-held base
    double* s_k = sEdgeM + nblk_max * ST_COLS;              // [8] conformal factor
    double* s_al = s_k + ST_COLS / 2;                       // [8] supertranslation
    double* s_xa = s_al + ST_COLS / 2;                      // [8] abscissa of the tile's first knot
    float* s_slope = reinterpret_cast<float*>(s_xa + ST_COLS / 2);   // [8] intervals per unit of x
    int* s_jlo = reinterpret_cast<int*>(s_slope + ST_COLS / 2);      // [8] first output of the column in this tile
    int* s_jhi = s_jlo + ST_COLS / 2;                       // [8] one past the last
    int* s_nch = s_jhi + ST_COLS / 2;                       // [8] chunks of 32 outputs
#define SY(r_, c_) sY[st_idx((r_) - ybase, (c_))]
#define SD(r_, c_) sD[st_idx((r_) - ybase, (c_))]
#define STAB(r_, f_) sTab[((r_) - ybase) * ST_TAB6 + (f_)]
#define STT(r_) sT[(r_) - ybase]

    // ---- stage the tile: thread 0 arms the barrier and issues the three asynchronous copies
    if (tid == 0) {
        const unsigned mb = st_smem_u32(&s_mbar);
        const int trows = ((yhi < N - 1) ? yhi : N - 1) - ybase + 1;
        const unsigned tab_bytes = (unsigned)trows * ST_TAB6 * (unsigned)sizeof(double);
        const unsigned box_bytes = (unsigned)box_rows * ST_ROWD * (unsigned)sizeof(double);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(2u * box_bytes + tab_bytes) : "memory");
        const int row0 = bz * N + ybase;                    // bz: series of a batch, stacked along the rows
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                    st_smem_u32(sY) + h2 * box_bytes),
                "l"(&tmF), "r"(col0), "r"(row0 + h2 * box_rows), "r"(mb)
                : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(st_smem_u32(sTab)),
                     "l"(tab + (size_t)ybase * ST_TAB6), "r"(tab_bytes), "r"(mb)
                     : "memory");
    }
    for (int e = tid; e <= yhi - ybase; e += nthr) sT[e] = t[ybase + e];
    // per-column constants and (MODE 0) the output range of each complex column, found while the copies fly
    if (tid < ST_COLS / 2) {
        const int c = tid;
        const int g = bx * (ST_COLS / 2) + c;
        const bool ok = (MODE == 0 && g < G);
        const double k = ok ? kconf[g] : 1.0, al = ok ? alpha[g] : 0.0;
        s_k[c] = k;
        s_al[c] = al;
        int jlo = 0, jhi = 0;
        if (ok) {                                           // interval guess of the evaluation: i ~ a + (u - x_a) * slope
            const double xa = __dmul_rn(k, __dsub_rn(t[a], al));
            const double xbv = __dmul_rn(k, __dsub_rn(t[b], al));
            s_xa[c] = xa;
            s_slope[c] = (float)((double)(b - a) / (xbv - xa));
            jlo = J[(size_t)by * G + g];
            jhi = J[(size_t)(by + 1) * G + g];
        }
        s_jlo[c] = jlo;
        s_jhi[c] = jhi;
        s_nch[c] = (jhi - jlo + 31) >> 5;                   // chunks of 32 outputs
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    // (MODE 0) evaluation work: chunks of 32 consecutive outputs of one column; the flattened (column, chunk) list is cut
    // into one contiguous range per warp, so a warp stays in one column for several chunks (its constants and its output
    // pointer are then loop invariants).  (ec, el) = first (column, chunk) of my range, erem = chunks left in it.
    int ec = 0, el = 0, erem = 0;
    double u_next = 0.0;
    if (MODE == 0) {
        int total = 0;
#pragma unroll
        for (int c = 0; c < ST_COLS / 2; ++c) total += s_nch[c];
        const int per = (total + nwarp - 1) / nwarp;
        el = warp * per;
        erem = (total - el < per) ? total - el : per;
        while (ec < ST_COLS / 2 && el >= s_nch[ec]) el -= s_nch[ec++];
        if (erem > 0 && ec < ST_COLS / 2) {               // output times of my first chunk: their latency hides behind the sweeps
            const int jn = s_jlo[ec] + (el << 5) + lane;
            if (jn < s_jhi[ec]) u_next = up[jn];
        }
    }
    {   // the tile has landed?
        unsigned ok = 0;
        const unsigned mb = st_smem_u32(&s_mbar);
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(mb) : "memory");
    }

    // ---- sweeps: half-warp = 16 real columns of one block of 16 rows; the block's d / m values stay in registers from the
    // forward sweep to the final store (one shared-memory load and one store per element)
    const int cc = tid & (ST_COLS - 1);
    const int q = qlo + tid / ST_COLS;
    const bool active = (q <= qhi);
    const int qbase = q * ST_BR;
    const int r0 = (qbase > slo) ? qbase : slo;
    const int r1 = (qbase + ST_BR - 1 < shi) ? qbase + ST_BR - 1 : shi;
    const bool full = active && r0 == qbase && r1 == qbase + ST_BR - 1;   // all but the blocks at the ends of the series
    // row qbase + k, my column: qbase - ybase is a multiple of 8, so the swizzle phase of step k is the constant k & 7
    const int rb = qbase - ybase;
    const int c2 = cc >> 1, c1 = cc & 1;
    const double* yq = sY + rb * ST_ROWD + c1;
    double* mq = sD + rb * ST_ROWD + c1;
    const double* tq = sTab + rb * ST_TAB6;
#define YQ(k_) yq[(k_) * ST_ROWD + ((c2 ^ ((k_) & 7)) << 1)]
#define MQ(k_) mq[(k_) * ST_ROWD + ((c2 ^ ((k_) & 7)) << 1)]
#define TQ2(k_, f_) (*reinterpret_cast<const double2*>(tq + (k_) * ST_TAB6 + (f_)))
    double dreg[ST_BR];
    if (full) {                                             // forward from a zero start
        double yc = YQ(0);
        double dym = yc - YQ(-1);
        double d = 0.0;
#pragma unroll
        for (int k = 0; k < ST_BR; ++k) {
            const double yn = YQ(k + 1);
            const double dy = yn - yc;
            const double2 pq = TQ2(k, 0);
            d = fma(-tq[k * ST_TAB6 + 4], d, pq.x * dy - pq.y * dym);
            dreg[k] = d;
            yc = yn;
            dym = dy;
        }
        sEdgeD[(q - qlo) * ST_COLS + cc] = d;
    } else if (active) {
        double yc = SY(r0, cc);
        double dym = yc - SY(r0 - 1, cc);
        double d = 0.0;
#pragma unroll
        for (int k = 0; k < ST_BR; ++k) {
            dreg[k] = 0.0;
            if (qbase + k >= r0 && qbase + k <= r1) {
                const double yn = YQ(k + 1);
                const double dy = yn - yc;
                const double2 pq = TQ2(k, 0);
                d = fma(-tq[k * ST_TAB6 + 4], d, pq.x * dy - pq.y * dym);
                dreg[k] = d;
                yc = yn;
                dym = dy;
            }
        }
        sEdgeD[(q - qlo) * ST_COLS + cc] = d;
    }
    __syncthreads();
    if (active) {
        // true d at the row before my block: block-end values of the blocks below, weighted by their Phi products
        double D = 0.0, w = 1.0;
        for (int p = q - 1; p >= qlo; --p) {
            D = fma(w, sEdgeD[(p - qlo) * ST_COLS + cc], D);
            w *= STAB(p * ST_BR + ST_BR - 1, 2);
            if (fabs(w) < 1e-24) break;
        }
        double m = 0.0;                                     // backward from a zero start, on the corrected d
        if (full) {
#pragma unroll
            for (int k = ST_BR - 1; k >= 0; --k) {
                const double2 pc = TQ2(k, 2);
                m = fma(-pc.y, m, fma(pc.x, D, dreg[k]));
                dreg[k] = m;
            }
        } else {
#pragma unroll
            for (int k = ST_BR - 1; k >= 0; --k)
                if (qbase + k >= r0 && qbase + k <= r1) {
                    const double2 pc = TQ2(k, 2);
                    m = fma(-pc.y, m, fma(pc.x, D, dreg[k]));
                    dreg[k] = m;
                }
        }
        sEdgeM[(q - qlo) * ST_COLS + cc] = m;               // m0 at my block start
    }
    __syncthreads();
    if (active) {
        double E = 0.0, w = 1.0;
        for (int p = q + 1; p <= qhi; ++p) {
            E = fma(w, sEdgeM[(p - qlo) * ST_COLS + cc], E);
            w *= STAB(p * ST_BR, 5);
            if (fabs(w) < 1e-24) break;
        }
        if (full) {
#pragma unroll
            for (int k = 0; k < ST_BR; ++k) MQ(k) = fma(tq[k * ST_TAB6 + 5], E, dreg[k]);
        } else {
#pragma unroll
            for (int k = 0; k < ST_BR; ++k)
                if (qbase + k >= r0 && qbase + k <= r1) MQ(k) = fma(tq[k * ST_TAB6 + 5], E, dreg[k]);
        }
    }
#undef YQ
#undef MQ
#undef TQ2
    __syncthreads();
    // end moments from the not-a-knot conditions
    if (a == 0 && tid < ST_COLS) {
        const double h0 = STT(1) - STT(0), h1 = STT(2) - STT(1);
        SD(0, tid) = ((h0 + h1) * SD(1, tid) - h0 * SD(2, tid)) / h1;
    }
    if (last_tile && tid >= 32 && tid < 32 + ST_COLS) {
        const int c = tid - 32;
        const double hm = STT(N - 2) - STT(N - 3), hp = STT(N - 1) - STT(N - 2);
        SD(N - 1, c) = ((hm + hp) * SD(N - 2, c) - hp * SD(N - 3, c)) / hm;
    }
    if (a == 0 || last_tile) __syncthreads();

    // ---- outputs
    if (MODE == 0) {
        double2* o2 = reinterpret_cast<double2*>(out);
        const int tmask = (1 << tshift) - 1;
        const int64_t jbase = (int64_t)bz * Nout;           // output rows of this series across the batch
        const int64_t ostride = ((int64_t)32 >> tshift) * ((int64_t)G << tshift);   // 32 outputs further down one column
        while (erem > 0 && ec < ST_COLS / 2) {
            const int c = ec;
            const int nrun = (s_nch[c] - el < erem) ? s_nch[c] - el : erem;       // my chunks in this column
            const int j1 = s_jhi[c];
            const int g = bx * (ST_COLS / 2) + c;
            const double k = s_k[c], al = s_al[c], xa = s_xa[c];
            const float slope = s_slope[c];
            int j = s_jlo[c] + (el << 5) + lane;
            const int64_t jg0 = jbase + j;
            double2* optr = o2 + ((((jg0 >> tshift) * G + g) << tshift) + (jg0 & tmask));   // [rows / tile, G, tile]; tile 1: time-major
            double u = u_next;
            for (int run = 0; run < nrun; ++run, j += 32, optr += ostride) {
                if (run + 1 < nrun) {                       // next chunk of the same column
                    u_next = (j + 32 < j1) ? up[j + 32] : 0.0;
                }
                if (j < j1) {
            int i = a + __float2int_rd((float)(u - xa) * slope);    // guess, verified below
            i = max(a, min(i, b - 1));
            double ti = STT(i), ti1 = STT(i + 1);
            double xi = __dmul_rn(k, __dsub_rn(ti, al));
            double xi1 = __dmul_rn(k, __dsub_rn(ti1, al));
            if (!((xi <= u || i == a) && (u < xi1 || i == b - 1))) {
                int lo_s = a, hi_s = b - 1;                 // largest i in [a, b-1] with x_i <= u (a if none)
                while (lo_s < hi_s) {
                    const int mid = (lo_s + hi_s + 1) >> 1;
                    if (__dmul_rn(k, __dsub_rn(STT(mid), al)) <= u) lo_s = mid; else hi_s = mid - 1;
                }
                i = lo_s;
                ti = STT(i);
                ti1 = STT(i + 1);
                xi = __dmul_rn(k, __dsub_rn(ti, al));
                xi1 = __dmul_rn(k, __dsub_rn(ti1, al));
            }
            const double ht = ti1 - ti;                     // h_i as the table formed it (same subtraction)
            const double hx = xi1 - xi;
            // 1/hx: two Newton steps from a single-precision reciprocal are exact to rounding; a true division only when
            // the step leaves the single-precision range
            float seed;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(seed) : "f"((float)hx));
            double inv_h = (double)seed;
            double e = fma(-hx, inv_h, 1.0);
            if (fabs(e) < 1e-3) {
                inv_h = fma(inv_h, e, inv_h);
                e = fma(-hx, inv_h, 1.0);
                inv_h = fma(inv_h, e, inv_h);
            } else {
                inv_h = 1.0 / hx;
            }
            const double A = (xi1 - u) * inv_h;
            const double B = (u - xi) * inv_h;
            const double h26 = ht * ht * (1.0 / 6.0);
            const double ca = (A * A * A - A) * h26;
            const double cb = (B * B * B - B) * h26;
            const int ir = i - ybase;
            const int o0 = ir * ST_ROWD + ((c ^ (ir & 7)) << 1), o1 = (ir + 1) * ST_ROWD + ((c ^ ((ir + 1) & 7)) << 1);
            const double2 yi = *reinterpret_cast<const double2*>(sY + o0);
            const double2 yi1 = *reinterpret_cast<const double2*>(sY + o1);
            const double2 Mi = *reinterpret_cast<const double2*>(sD + o0);
            const double2 Mi1 = *reinterpret_cast<const double2*>(sD + o1);
            double2 r;
            r.x = A * yi.x + B * yi1.x + (ca * Mi.x + cb * Mi1.x);
            r.y = A * yi.y + B * yi1.y + (ca * Mi.y + cb * Mi1.y);
                    *optr = r;
                }
                u = u_next;
            }
            erem -= nrun;
            el = 0;
            ++ec;
            if (erem > 0 && ec < ST_COLS / 2) {             // first chunk of the next column
                const int jn = s_jlo[ec] + lane;
                u_next = (jn < s_jhi[ec]) ? up[jn] : 0.0;
            }
        }
    } else {
        // one thread per (knot, real column): 128-byte row segments of the time-major output
        const int rstep = nthr / ST_COLS;
        const bool colok = col0 + cc < G2;
        out += (size_t)bz * N * G2;
        if (MODE == 4) up += (size_t)bz * N * G2;
        auto value = [&](int i, bool& last_row, double& last_val) -> double {
            const double ht = STT(i + 1) - STT(i);
            const double yi = SY(i, cc), yi1 = SY(i + 1, cc);
            const double Mi = SD(i, cc), Mi1 = SD(i + 1, cc);
            last_row = false;
            if (MODE == 1) {
                const double dl = (yi1 - yi) / ht, h6 = ht * (1.0 / 6.0);
                if (i == N - 2) {
                    last_row = true;
                    last_val = dl + h6 * (Mi + 2.0 * Mi1);
                }
                return dl - h6 * (2.0 * Mi + Mi1);
            } else if (MODE == 2) {
                if (i == N - 2) {
                    last_row = true;
                    last_val = Mi1;
                }
                return Mi;
            } else if (MODE == 3) {
                return 0.5 * ht * (yi + yi1) - (ht * ht * ht) * (1.0 / 24.0) * (Mi + Mi1);
            } else {
                // second antiderivative over the interval: II_{i+1} - II_i = h I_i + int int of the cubic piece
                const double c1 = (yi1 - yi) / ht - ht * (1.0 / 6.0) * (2.0 * Mi + Mi1);
                const double h2 = ht * ht;
                const double Jv = h2 * (0.5 * yi + ht * (1.0 / 6.0) * c1 + h2 * ((1.0 / 24.0) * Mi + (1.0 / 120.0) * (Mi1 - Mi)));
                return Jv + (colok ? ht * up[(size_t)i * G2 + col0 + cc] : 0.0);
            }
        };
        if (MODE <= 2) {
            for (int i = a + tid / ST_COLS; i < b; i += rstep) {
                bool lr;
                double lv = 0.0;
                const double v = value(i, lr, lv);
                if (!colok) continue;
                out[(size_t)i * G2 + col0 + cc] = v;
                if (lr) out[(size_t)(N - 1) * G2 + col0 + cc] = lv;
            }
        } else {
            // antiderivatives: the increments of the tile's intervals are summed up inside the tile (row i + 1 receives the
            // sum over the intervals a .. i), so the column scan that follows only has to carry one total per tile across
            // the tiles - one read-modify-write pass over the output less than scanning the raw increments
            constexpr int VMAX = 16;                        // rows per thread: body / (threads / 16) < 16
            double val[VMAX];
            bool lr;
            double lv;
#pragma unroll
            for (int k = 0; k < VMAX; ++k) {
                const int i = a + tid / ST_COLS + k * rstep;
                val[k] = (i < b) ? value(i, lr, lv) : 0.0;
            }
            __syncthreads();                                // every moment has been read: the M tile becomes the work area
#pragma unroll
            for (int k = 0; k < VMAX; ++k) {
                const int i = a + tid / ST_COLS + k * rstep;
                if (i < b) SD(i, cc) = val[k];
            }
            __syncthreads();
            const int nr = b - a, slen = (nr + ST_BR - 1) / ST_BR;      // 16 strips of slen rows per column
            for (int id = tid; id < ST_BR * ST_COLS; id += nthr) {
                const int s = id / ST_COLS, c = id - s * ST_COLS;
                const int r0s = a + s * slen, r1s = (r0s + slen < b) ? r0s + slen : b;
                double run = 0.0;
                for (int r = r0s; r < r1s; ++r) {
                    run += SD(r, c);
                    SD(r, c) = run;
                }
                sEdgeD[s * ST_COLS + c] = run;
            }
            __syncthreads();
            if (tid < ST_COLS) {                            // strip totals -> what precedes each strip
                double run = 0.0;
#pragma unroll
                for (int p = 0; p < ST_BR; ++p) {
                    const double v = sEdgeD[p * ST_COLS + tid];
                    sEdgeD[p * ST_COLS + tid] = run;
                    run += v;
                }
            }
            __syncthreads();
            if (colok) {
                const float rs = 1.0f / (float)slen;
                for (int i = a + tid / ST_COLS; i < b; i += rstep) {
                    int s = (int)((float)(i - a) * rs);    // (i - a) / slen: both below 2^9, one correction step at most
                    s -= (s * slen > i - a);
                    s += ((s + 1) * slen <= i - a);
                    out[(size_t)(i + 1) * G2 + col0 + cc] = SD(i, cc) + sEdgeD[s * ST_COLS + cc];
                    if (i == 0) out[col0 + cc] = 0.0;
                }
            }
        }
    }
#undef SY
#undef SD
#undef STAB
#undef STT
}

// Column scan that turns the per-interval integrals into the antiderivative (waveform_base.py:697-703), in place and
// without a workspace.  After spline_tile_kernel<3 / 4>: rows k body + 1 .. min((k + 1) body, N - 1) of x hold the running sums of tile k; its last
// row holds the tile total.  (B) one thread per column turns the tile totals into global values (N / body steps); (C) every
// tile but the first adds the (now global) value at the end of the previous tile to all its rows but the last.
__global__ void __launch_bounds__(1024)
tile_scan_totals_kernel(double* __restrict__ x, int64_t N, int C, int body, int ntiles) {
    // 32 columns x 32 segments of the tile range per CTA: a serial walk over N / body totals per column is a chain of
    // dependent DRAM round trips (4 ms for 1e6 rows); here every thread walks ntiles / 32 of them with 8 loads in flight
    __shared__ double tot[32][33];
    const int col = blockIdx.x * 32 + threadIdx.x, seg = threadIdx.y;
    const int per = (ntiles + 31) / 32;
    const int k0 = seg * per, k1 = (k0 + per < ntiles) ? k0 + per : ntiles;
    const bool ok = col < C;
    auto at = [&](int k) -> double* {
        const int64_t last = (int64_t)(k + 1) * body;
        return x + (last < N - 1 ? last : N - 1) * C + col;
    };
    double run = 0.0;
    if (ok) {
        for (int k = k0; k < k1; k += 8) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (k + j < k1) ? *at(k + j) : 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (k + j < k1) {
                    run += v[j];
                    *at(k + j) = run;
                }
        }
    }
    tot[seg][threadIdx.x] = run;
    __syncthreads();
    if (!ok || seg == 0) return;
    double pre = 0.0;
    for (int p = 0; p < seg; ++p) pre += tot[p][threadIdx.x];
    for (int k = k0; k < k1; k += 8) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (k + j < k1) ? *at(k + j) : 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (k + j < k1) *at(k + j) = v[j] + pre;
    }
}

__global__ void __launch_bounds__(128)
tile_scan_add_kernel(double* __restrict__ x, int64_t N, int C, int body) {
    const int col = blockIdx.x * 128 + threadIdx.x;
    if (col >= C) return;
    const int64_t first = (int64_t)(blockIdx.y + 1) * body;             // last row of the previous tile: already global
    if (first >= N - 1) return;
    const int64_t last = first + body < N - 1 ? first + body : N - 1;   // my own last row: already global
    const double base = x[first * C + col];
    double* p = x + (first + 1) * C + col;
    int left = (int)(last - first - 1);
    for (; left >= 8; left -= 8, p += 8 * (size_t)C) {       // 8 loads in flight before the first store
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = p[(size_t)j * C];
#pragma unroll
        for (int j = 0; j < 8; ++j) p[(size_t)j * C] = v[j] + base;
    }
    for (; left > 0; --left, p += C) *p += base;
}

static int launch_tile_scan(double* x, int64_t N, int C, int body, cudaStream_t st) {
    const unsigned gx = (unsigned)((C + 127) / 128);
    const int64_t ntiles = (N - 1 + body - 1) / body;
    if (ntiles <= 1) return SCRIB200_OK;
    tile_scan_totals_kernel<<<(unsigned)((C + 31) / 32), dim3(32, 32), 0, st>>>(x, N, C, body, (int)ntiles);
    SCRIB200_CHECK_LAUNCH("spline_calculus(scan: totals)");
    tile_scan_add_kernel<<<dim3(gx, (unsigned)(ntiles - 1)), 128, 0, st>>>(x, N, C, body);
    SCRIB200_CHECK_LAUNCH("spline_calculus(scan: add)");
    return SCRIB200_OK;
}

static int tile_box_rows(int body, int halo) {
    const int ymax = body + 2 * halo + 2;                    // rows a tile needs; + 7 for the 8-row alignment of its first row
    return (((ymax + 7 + 1) / 2) + 7) & ~7;
}

static size_t tile_smem_bytes(int body, int halo) {
    const size_t nrows = 2 * (size_t)tile_box_rows(body, halo);
    const size_t nblk = std::max<size_t>(((size_t)body + 2 * halo) / ST_BR, ST_BR);
    return (2 * nrows * ST_ROWD + nrows * ST_TAB6 + nrows + 2 * nblk * ST_COLS + 5 * ST_COLS) * sizeof(double) + 1024;   // + constants, alignment slack
}

static void resolve_tile(int& halo, int& body) {
    if (halo <= 0) halo = 32;
    if (body <= 0) {
        static const int env_body = [] { const char* e = std::getenv("SCRIB200_TILE_BODY"); return e ? std::atoi(e) : 0; }();
        body = env_body > 0 ? env_body : ((halo <= 32) ? 240 : (halo <= 64 ? 176 : 128));   // two CTAs per SM up to halo = 64
    }
}

typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TensorMapEncodeTiledFn tensor_map_encoder() {
    static TensorMapEncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TensorMapEncodeTiledFn>(p);
    }
    return fn;
}

template <int MODE>
static int launch_tile(const double* t, int64_t n_times, const double* F, int G, const double* kconf, const double* alpha,
                       const double* tab, const double* uprm, int64_t n_out, double* out, int tshift, int halo, int body,
                       void* workspace, size_t workspace_bytes, void* stream, const char* name, int n_series = 1,
                       int64_t row_lo = 0, int64_t row_hi = -1) {
    SCRIB200_REQUIRE(n_times < (int64_t)2147483000 && n_out < (int64_t)2147483000, "%s: series longer than 2^31 samples", name);
    resolve_tile(halo, body);
    SCRIB200_REQUIRE(halo % ST_BR == 0 && body % ST_BR == 0, "%s: body=%d and halo=%d must be multiples of %d", name, body, halo, ST_BR);
    const size_t smem = tile_smem_bytes(body, halo);
    const int nblk = (body + 2 * halo) / ST_BR;
    const int box_rows = tile_box_rows(body, halo);
    SCRIB200_REQUIRE(nblk <= ST_MAXBLK && smem <= 220 * 1024 && box_rows <= 256, "%s: body=%d halo=%d does not fit one CTA", name, body, halo);
    const int64_t ntiles = (n_times - 1 + body - 1) / body;
    SCRIB200_REQUIRE(ntiles <= 65535, "%s: too many time tiles (%lld); raise `body`", name, (long long)ntiles);
    SCRIB200_REQUIRE(n_series >= 1 && n_series <= 65535 && (int64_t)n_series * n_times < (int64_t)2147483000,
                     "%s: n_series=%d (x %lld rows) out of range", name, n_series, (long long)n_times);
    // F as a [n_series * n_times, 2G] FP64 tensor; box = one tile half: 16 columns x box_rows rows, 128-byte swizzle
    TensorMapEncodeTiledFn encode = tensor_map_encoder();
    SCRIB200_REQUIRE(encode != nullptr, "%s: cuTensorMapEncodeTiled is not available from this driver", name);
    CUtensorMap tmF;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)(2 * (int64_t)G), (cuuint64_t)((int64_t)n_series * n_times)};
        const cuuint64_t gstride[1] = {(cuuint64_t)(2 * (int64_t)G) * sizeof(double)};
        const cuuint32_t box[2] = {(cuuint32_t)ST_COLS, (cuuint32_t)box_rows};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode(&tmF, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(F), gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SCRIB200_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed (%d) for a [%lld, %d] grid array", name, (int)r,
                         (long long)n_series * n_times, 2 * G);
    }
    const int threads = ((nblk * ST_COLS + 31) / 32) * 32;   // one half-warp per sweep block
    const bool small = threads <= 320;                       // default tiles (19 blocks): 102 registers per thread at 2 CTAs / SM
    if (small) cudaFuncSetAttribute(spline_tile_kernel<MODE, 320>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else cudaFuncSetAttribute(spline_tile_kernel<MODE, ST_MAXBLK * ST_COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // tiles to launch: all of them, or only those that hold input rows [row_lo, row_hi) (a caller that evaluates a slice of
    // the output times and knows which input rows it can touch)
    int64_t tile_lo = 0, tile_hi = ntiles;
    if (row_hi >= 0) {
        tile_lo = (row_lo > 0 ? row_lo : 0) / body;
        tile_hi = (row_hi + body - 1) / body;
        if (tile_hi > ntiles) tile_hi = ntiles;
        if (tile_lo >= tile_hi) return SCRIB200_OK;
    }
    dim3 grid((2 * G + ST_COLS - 1) / ST_COLS, (unsigned)(tile_hi - tile_lo), (unsigned)n_series);
    int* J = nullptr;
    int* flags = nullptr;
    if (MODE == 0) {
        const size_t need = ((size_t)(ntiles + 1) * G + (size_t)ntiles * grid.x) * sizeof(int);   // tables cover every tile
        SCRIB200_REQUIRE(workspace && workspace_bytes >= need, "%s: workspace too small (%zu < %zu)", name, workspace_bytes, need);
        J = reinterpret_cast<int*>(workspace);
        const int64_t total = (ntiles + 1) * (int64_t)G;
        spline_jrange_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            t, (int)n_times, G, kconf, alpha, uprm, (int)n_out, body, (int)ntiles, J);
        SCRIB200_CHECK_LAUNCH(name);
        flags = J + total;
        const int64_t nflags = ntiles * (int64_t)grid.x;
        spline_tile_flags_kernel<<<(unsigned)((nflags + 255) / 256), 256, 0, (cudaStream_t)stream>>>(J, (int)ntiles, G, (int)grid.x, flags);
        SCRIB200_CHECK_LAUNCH(name);
    }
    if (small)
        spline_tile_kernel<MODE, 320><<<grid, threads, smem, (cudaStream_t)stream>>>(
            tmF, t, (int)n_times, G, kconf, alpha, tab, uprm, (int)n_out, out, tshift, body, halo, J, flags, box_rows, (int)tile_lo);
    else
        spline_tile_kernel<MODE, ST_MAXBLK * ST_COLS><<<grid, threads, smem, (cudaStream_t)stream>>>(
            tmF, t, (int)n_times, G, kconf, alpha, tab, uprm, (int)n_out, out, tshift, body, halo, J, flags, box_rows, (int)tile_lo);
    SCRIB200_CHECK_LAUNCH(name);
    return SCRIB200_OK;
}

}  // namespace scrib200

extern "C" int scrib200_spline_prepare(const double* t, int64_t n_times, double gamma_factor, int divide, double time_translation,
                                       const double* kconf, const double* alpha, int G, double* tab, double* uprm,
                                       double* info, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(t && tab && info, "spline_prepare: null pointer");
    SCRIB200_REQUIRE(n_times >= 4 && n_times < (int64_t)2147483000, "spline_prepare: a cubic interpolating spline needs at least 4 knots; got %lld",
                     (long long)n_times);
    SCRIB200_REQUIRE(uprm == nullptr || (kconf && alpha && G > 0), "spline_prepare: output times need kconf, alpha, G");
    SCRIB200_REQUIRE(aligned16(tab), "spline_prepare: tab must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    // seed info in stream order (lo = hi = N means "empty / not found")
    spline_info_init_kernel<<<1, 32, 0, st>>>(info, (double)n_times);
    const unsigned blocks = (unsigned)((n_times + 255) / 256);
    spline_factor_kernel<<<blocks, 256, 0, st>>>(t, (int)n_times, tab, gamma_factor, divide,
                                                 time_translation, kconf, alpha, G, uprm, info);
    SCRIB200_CHECK_LAUNCH("spline_prepare(factor)");
    spline_decay_kernel<<<blocks, 256, 0, st>>>(tab, (int)n_times, info);
    SCRIB200_CHECK_LAUNCH("spline_prepare(decay)");
    return SCRIB200_OK;
}

extern "C" size_t scrib200_spline_remap_workspace_bytes(int64_t n_times, int G, int halo, int body) {
    using namespace scrib200;
    resolve_tile(halo, body);
    if (n_times < 2 || body <= 0) return 16;
    const int64_t ntiles = (n_times - 1 + body - 1) / body;
    const size_t ncg = (size_t)(2 * G + ST_COLS - 1) / ST_COLS;
    return ((size_t)(ntiles + 1) * (size_t)G + (size_t)ntiles * ncg) * sizeof(int) + 16;
}

extern "C" int scrib200_spline_remap(const double* t, int64_t n_times, const double* F, int G, const double* kconf,
                                     const double* alpha, const double* tab, const double* uprm, int64_t n_out,
                                     double* out, int tile, int halo, int body, int n_series, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(t && F && kconf && alpha && tab && uprm && out, "spline_remap: null pointer");
    SCRIB200_REQUIRE(n_times >= 4, "spline_remap: a cubic interpolating spline needs at least 4 knots; got %lld",
                     (long long)n_times);
    int tshift = 0;
    while ((1 << tshift) < tile) ++tshift;
    SCRIB200_REQUIRE(G > 0 && (tile == 0 || (tile >= 2 && tile <= 32 && (1 << tshift) == tile)),
                     "spline_remap: G=%d, tile=%d must be 0 (time-major output) or a power of two in 2..32", G, tile);
    SCRIB200_REQUIRE(aligned16(F) && aligned16(out) && aligned16(tab), "spline_remap: pointers must be 16-byte aligned");
    if (n_out <= 0) return SCRIB200_OK;
    return launch_tile<0>(t, n_times, F, G, kconf, alpha, tab, uprm, n_out, out, tile ? tshift : 0, halo, body, workspace,
                          workspace_bytes, stream, "spline_remap", n_series);
}

extern "C" int scrib200_spline_remap_rows(const double* t, int64_t n_times, const double* F, int G, const double* kconf,
                                          const double* alpha, const double* tab, const double* uprm, int64_t n_out,
                                          double* out, int tile, int halo, int body, int n_series, int64_t row_lo,
                                          int64_t row_hi, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(t && F && kconf && alpha && tab && uprm && out, "spline_remap_rows: null pointer");
    SCRIB200_REQUIRE(n_times >= 4, "spline_remap_rows: a cubic interpolating spline needs at least 4 knots; got %lld", (long long)n_times);
    int tshift = 0;
    while ((1 << tshift) < tile) ++tshift;
    SCRIB200_REQUIRE(G > 0 && (tile == 0 || (tile >= 2 && tile <= 32 && (1 << tshift) == tile)),
                     "spline_remap_rows: G=%d, tile=%d must be 0 (time-major output) or a power of two in 2..32", G, tile);
    SCRIB200_REQUIRE(aligned16(F) && aligned16(out) && aligned16(tab), "spline_remap_rows: pointers must be 16-byte aligned");
    SCRIB200_REQUIRE(row_lo >= 0 && row_hi >= row_lo, "spline_remap_rows: rows [%lld, %lld)", (long long)row_lo, (long long)row_hi);
    if (n_out <= 0) return SCRIB200_OK;
    return launch_tile<0>(t, n_times, F, G, kconf, alpha, tab, uprm, n_out, out, tile ? tshift : 0, halo, body, workspace,
                          workspace_bytes, stream, "spline_remap_rows", n_series, row_lo, row_hi);
}

extern "C" int scrib200_spline_calculus(const double* t, int64_t n_times, const double* data, int ncol,
                                        const double* tab, int order, double* out, double* aux, int halo, int body,
                                        void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(t && data && tab && out, "spline_calculus: null pointer");
    SCRIB200_REQUIRE(n_times >= 4, "spline_calculus: a cubic spline needs at least 4 knots; got %lld", (long long)n_times);
    SCRIB200_REQUIRE(order == 1 || order == 2 || order == -1 || order == -2,
                     "spline_calculus: order must be 1, 2 (derivatives) or -1, -2 (antiderivatives)");
    SCRIB200_REQUIRE(order != -2 || aux, "spline_calculus: order -2 needs `aux` [n_times, ncol] (receives the first antiderivative)");
    SCRIB200_REQUIRE(ncol > 0, "spline_calculus: ncol=%d", ncol);
    SCRIB200_REQUIRE(aligned16(data) && aligned16(out) && aligned16(tab) && aligned16(aux), "spline_calculus: pointers must be 16-byte aligned");
    if (order == 1)
        return launch_tile<1>(t, n_times, data, ncol, nullptr, nullptr, tab, nullptr, 0, out, 0, halo, body, nullptr, 0, stream, "spline_calculus");
    if (order == 2)
        return launch_tile<2>(t, n_times, data, ncol, nullptr, nullptr, tab, nullptr, 0, out, 0, halo, body, nullptr, 0, stream, "spline_calculus");
    const int C = 2 * ncol;
    double* first = (order == -1) ? out : aux;
    int rbody = body, rhalo = halo;
    resolve_tile(rhalo, rbody);                              // the tile height the launches below will use
    int rc = launch_tile<3>(t, n_times, data, ncol, nullptr, nullptr, tab, nullptr, 0, first, 0, halo, body, nullptr, 0, stream, "spline_calculus");
    if (rc != SCRIB200_OK) return rc;
    rc = launch_tile_scan(first, n_times, C, rbody, (cudaStream_t)stream);
    if (rc != SCRIB200_OK) return rc;
    if (order == -1) return SCRIB200_OK;
    rc = launch_tile<4>(t, n_times, data, ncol, nullptr, nullptr, tab, first, 0, out, 0, halo, body, nullptr, 0, stream, "spline_calculus");
    if (rc != SCRIB200_OK) return rc;
    return launch_tile_scan(out, n_times, C, rbody, (cudaStream_t)stream);
}
