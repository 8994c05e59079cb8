// K2 / K8 (second generation): batched not-a-knot cubic splines along time, tile-resident in shared memory.
//
// Replaces scri/waveform_grid.py:576-588 (two scipy InterpolatedUnivariateSpline objects per grid point on the knots
// x_i = k[g] (t[i] - alpha[g]), evaluated at the common output times u') and scri/waveform_base.py:689-703
// (CubicSpline(t, data).derivative(k)(t), .antiderivative(k)(t)).
//
// Two observations shape the kernel:
//  1. The knots of every grid point are an affine image of the same sample times t, and the interpolating cubic
//     spline is affine-invariant.  In t-units the moment system  T(t) M = 6 [y_{i-1}, y_i, y_{i+1}]  has ONE matrix
//     for all columns, so its Thomas factorisation is done once (`spline_factor_kernel`, one thread per row, the
//     c' recurrence contracts by <= 1/12 per row so 24 rows of run-in reproduce it to rounding) and stored as a
//     table of (P, Q, W, c') per row.  The per-column work is then division free:
//         d_i = P_i (y_{i+1}-y_i) - Q_i (y_i-y_{i-1}) - W_i d_{i-1}        (forward)
//         M_i = d_i - c'_i M_{i+1}                                         (backward)
//     The evaluation still uses the reference's own rounded abscissae x_i = fl(k fl(t_i - alpha)) for the position
//     inside the interval (one ulp of x at t ~ 1e4 moves the interpolant at the 1e-13 level); only the small
//     curvature term uses the shared factorisation (h_t^2 M_t = h_x^2 M_x up to 1e-11 relative, i.e. < 1e-14 of the
//     value).
//  2. Both recurrences forget their start geometrically (|W| ~ 0.27 on uniform samples), so a tile of `body`
//     intervals plus `halo` rows on each side is self-contained.  A CTA owns [body + 2 halo + 3] rows x 16 real
//     columns (8 complex grid points = 128-byte rows) of F in shared memory, read from HBM exactly once
//     (cp.async, 16 B per copy), sweeps them with half-warps (16 columns each) spread over sub-ranges of rows,
//     and evaluates the outputs that fall into its intervals with lanes running along the output index, so the
//     stores into the time-tiled output are full 128-byte runs.  No workspace, no checkpoints, F is never re-read.
//
// HBM traffic per (knot, complex column): 16 B x (1 + (2 halo + 3)/body) read + 16 B written.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"

namespace scrib200 {

constexpr int ST_COLS = 16;            // real columns per CTA
constexpr int ST_PITCH = ST_COLS + 2;  // doubles per shared-memory row (144 B: consecutive rows shift by 4 banks)
constexpr int FACTOR_RUNIN = 24;

__device__ __forceinline__ void st_cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void st_cp_async_commit_wait() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

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

// tab[i] = (P_i, Q_i, W_i, c'_i) for rows 1..N-2 (rows 0 and N-1 zero); also u'_i = inv_gamma (t_i - tt) and the
// retained output block [lo, hi) = { i : umin <= u'_i <= umax } (waveform_grid.py:564-568), umin/umax reduced here
// from k (t_0 - alpha), k (t_{N-1} - alpha) over the G grid points.
__global__ void __launch_bounds__(256)
spline_factor_kernel(const double* __restrict__ t, int N, double4* __restrict__ tab, double inv_gamma, double tt,
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
            const double u = __dmul_rn(inv_gamma, __dsub_rn(t[i], tt));
            uprm[i] = u;
            // u' is non-decreasing: exactly one i starts / ends the retained block
            const double up_prev = (i > 0) ? __dmul_rn(inv_gamma, __dsub_rn(t[i - 1], tt)) : -CUDART_INF;
            if (u >= umin && !(up_prev >= umin)) info[0] = (double)i;          // lo: first u' >= umin
            if (u > umax && !(up_prev > umax)) info[1] = (double)i;            // hi: first u' >  umax
            if (i == 0) {
                info[4] = umin;
                info[5] = umax;
            }
        }
    }
    if (i >= N) return;
    double4 row = make_double4(0.0, 0.0, 0.0, 0.0);
    if (i >= 1 && i <= N - 2) {
        int r = i - FACTOR_RUNIN;
        if (r < 1) r = 1;
        double cp = 0.0;
        for (; r <= i; ++r) {
            double sub, diag, sup, hm, hp;
            nak_row(t, N, r, sub, diag, sup, hm, hp);
            const double e = 1.0 / (diag - sub * cp);
            cp = sup * e;
            if (r == i) row = make_double4(6.0 * e / hp, 6.0 * e / hm, sub * e, cp);
        }
    }
    tab[i] = row;
}

__global__ void spline_info_init_kernel(double* __restrict__ info, double n) {
    if (threadIdx.x < 8) info[threadIdx.x] = (threadIdx.x < 2) ? n : 0.0;
}

// Worst decay of the two recurrences over any window of 32 / 64 consecutive rows:
// info[2] = max_i prod_{r=i-31..i} |W_r| (forward) or prod |c'_r| (backward), info[3] the same for 64 rows.
__global__ void __launch_bounds__(256)
spline_decay_kernel(const double4* __restrict__ tab, int N, double* __restrict__ info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double d32 = 0.0, d64 = 0.0;
    if (i >= 64 && i <= N - 2) {
        double pw = 1.0, pc = 1.0;
        for (int r = i; r > i - 64; --r) {
            const double4 v = tab[r];
            pw *= fabs(v.z);
            pc *= fabs(v.w);
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

// MODE 0: evaluate at up[] (the BMS remap); MODE 1 / 2: first / second derivative at the knots; MODE 3: the
// definite integral over [t_i, t_{i+1}] stored at row i+1 (row 0 = 0) - a column scan turns it into the antiderivative;
// MODE 4: the increment of the second antiderivative, `up` then holding the (scanned) first antiderivative [N, G].
// For MODE >= 1 `out` is [N, G] complex time-major and Nout/tshift are unused.
template <int MODE>
__global__ void __launch_bounds__(256, 2)
spline_tile_kernel(const double* __restrict__ t, int N, const double* __restrict__ F, int G,
                   const double* __restrict__ kconf, const double* __restrict__ alpha,
                   const double4* __restrict__ tab, const double* __restrict__ up, int Nout,
                   double* __restrict__ out, int tshift, int body, int halo) {
    extern __shared__ __align__(16) double s_mem[];
    __shared__ double s_k[ST_COLS / 2], s_al[ST_COLS / 2];
    __shared__ int s_jlo[ST_COLS / 2], s_jhi[ST_COLS / 2];

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int G2 = 2 * G;
    const int col0 = blockIdx.x * ST_COLS;                  // first real column of this CTA
    const int a = blockIdx.y * body;                        // intervals a .. b-1, knots a .. b
    const int b = (a + body < N - 1) ? a + body : N - 1;
    const bool last_tile = (b == N - 1);
    const int mhi = (b + halo < N - 2) ? b + halo : N - 2;  // last row of the backward sweep
    const int dlo = (a - 2 > 1) ? a - 2 : 1;                // first row whose d is kept (two below a: the right-end formula)
    const int flo = (a - halo > 1) ? a - halo : 1;          // first row of the forward sweep
    const int ylo = flo - 1, yhi = mhi + 1;                 // rows of F resident in shared memory
    const int nyrows = yhi - ylo + 1;
    const int ymax = body + 2 * halo + 3, dmax = body + halo + 4;

    double* sY = s_mem;                                     // [ymax][ST_PITCH]     row r -> r - ylo
    double* sD = sY + (size_t)ymax * ST_PITCH;              // [dmax][ST_PITCH]     row r -> r - a + 2   (d, then M)
    double4* sTab = reinterpret_cast<double4*>(sD + (size_t)dmax * ST_PITCH);   // [ymax]   row r -> r - ylo
    double* sT = reinterpret_cast<double*>(sTab + ymax);    // [ymax]               row r -> r - ylo
#define SY(r_, c_) sY[((r_) - ylo) * ST_PITCH + (c_)]
#define SD(r_, c_) sD[((r_) - a + 2) * ST_PITCH + (c_)]
#define STAB(r_) sTab[(r_) - ylo]
#define STT(r_) sT[(r_) - ylo]

    // ---- stage the tile: F rows (16-byte copies, 8 per row), the factor table rows and the sample times
    {
        const int nchunk = nyrows * (ST_COLS / 2);
        for (int e = tid; e < nchunk; e += nthr) {
            const int rr = e >> 3, ch = e & 7;
            double* dst = sY + rr * ST_PITCH + 2 * ch;
            if (col0 + 2 * ch < G2) {
                st_cp_async16(dst, F + (size_t)(ylo + rr) * G2 + col0 + 2 * ch);
            } else {
                dst[0] = 0.0;
                dst[1] = 0.0;
            }
        }
        for (int e = tid; e < nyrows; e += nthr) {
            st_cp_async16(&sTab[e], &tab[ylo + e]);
            st_cp_async16(reinterpret_cast<char*>(&sTab[e]) + 16, reinterpret_cast<const char*>(&tab[ylo + e]) + 16);
            sT[e] = t[ylo + e];
        }
    }
    // per-column constants and (MODE 0) the output range of each complex column, found while the copies fly
    if (tid < ST_COLS) {
        const int c = tid >> 1, which = tid & 1;
        const int g = blockIdx.x * (ST_COLS / 2) + c;
        const double k = (MODE == 0 && g < G) ? kconf[g] : 1.0, al = (MODE == 0 && g < G) ? alpha[g] : 0.0;
        if (which == 0) {
            s_k[c] = k;
            s_al[c] = al;
        }
        if (MODE == 0) {
            int j = (which == 0) ? 0 : Nout;
            const bool search = (which == 0) ? (a > 0) : !last_tile;
            if (search && g < G) {
                const double xb = __dmul_rn(k, __dsub_rn(t[which == 0 ? a : b], al));
                int lo_s = 0, hi_s = Nout;     // first j with up[j] >= xb
                while (lo_s < hi_s) {
                    const int mid = (lo_s + hi_s) >> 1;
                    if (up[mid] < xb) lo_s = mid + 1; else hi_s = mid;
                }
                j = lo_s;
            }
            if (which == 0) s_jlo[c] = j; else s_jhi[c] = j;
        }
    }
    st_cp_async_commit_wait();
    __syncthreads();

    // ---- sweeps: half-warp = 16 real columns of one sub-range of rows
    const int cc = tid & (ST_COLS - 1);
    const int nsub = nthr / ST_COLS;
    const int sidx = tid / ST_COLS;
    const int nrows = mhi - dlo + 1;                        // rows dlo .. mhi get a moment
    const int SR = (nrows + nsub - 1) / nsub;
    const int rs = dlo + sidx * SR;
    const int re = (rs + SR < mhi + 1) ? rs + SR : mhi + 1; // my rows: [rs, re)
    if (rs < re) {
        int r = (rs - halo > flo) ? rs - halo : flo;
        double yc = SY(r, cc);
        double dym = yc - SY(r - 1, cc);
        double d = 0.0;
        for (; r < rs; ++r) {                               // run-in: nothing stored
            const double yn = SY(r + 1, cc);
            const double dy = yn - yc;
            const double4 tb = STAB(r);
            d = fma(-tb.z, d, tb.x * dy - tb.y * dym);
            yc = yn;
            dym = dy;
        }
        for (; r < re; ++r) {
            const double yn = SY(r + 1, cc);
            const double dy = yn - yc;
            const double4 tb = STAB(r);
            d = fma(-tb.z, d, tb.x * dy - tb.y * dym);
            SD(r, cc) = d;
            yc = yn;
            dym = dy;
        }
    }
    __syncthreads();
    double M = 0.0;
    if (rs < re) {                                          // backward run-in over the rows above mine (read only)
        int r = (re + halo - 1 < mhi) ? re + halo - 1 : mhi;
        for (; r >= re; --r) M = fma(-STAB(r).w, M, SD(r, cc));
    }
    __syncthreads();
    if (rs < re) {
        for (int r = re - 1; r >= rs; --r) {
            M = fma(-STAB(r).w, M, SD(r, cc));
            SD(r, cc) = M;
        }
    }
    __syncthreads();
    // end moments from the not-a-knot conditions
    if (a == 0 && tid < ST_COLS) {
        const double h0 = STT(1) - STT(0), h1 = STT(2) - STT(1);
        SD(0, tid) = ((h0 + h1) * SD(1, tid) - h0 * SD(2, tid)) / h1;
    }
    if (last_tile && tid >= 32 && tid < 32 + ST_COLS) {
        const int c = tid - 32;
        const double hm = STT(N - 2) - STT(N - 3), hp = STT(N - 1) - STT(N - 2);
        // N == 4: M_2 and M_1 are both final; the formula reads M_{N-2}, M_{N-3}
        SD(N - 1, c) = ((hm + hp) * SD(N - 2, c) - hp * SD(N - 3, c)) / hm;
    }
    if (a == 0 || last_tile) __syncthreads();

    // ---- outputs
    if (MODE == 0) {
        const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
        double2* o2 = reinterpret_cast<double2*>(out);
        const int tmask = (1 << tshift) - 1;
        const int64_t tileGT = (int64_t)G << tshift;
        for (int c = warp; c < ST_COLS / 2; c += nwarp) {
            const int g = blockIdx.x * (ST_COLS / 2) + c;
            if (g >= G) continue;
            const double k = s_k[c], al = s_al[c];
            const int jlo = s_jlo[c], jhi = s_jhi[c];
            const double xa = __dmul_rn(k, __dsub_rn(STT(a), al));
            const double xbv = __dmul_rn(k, __dsub_rn(STT(b), al));
            const double slope = (double)(b - a) / (xbv - xa);
            for (int j = jlo + lane; j < jhi; j += 32) {
                const double u = up[j];
                int i = a + (int)fmin(fmax((u - xa) * slope, 0.0), (double)(b - 1 - a));
                double xi = __dmul_rn(k, __dsub_rn(STT(i), al));
                double xi1 = __dmul_rn(k, __dsub_rn(STT(i + 1), al));
                if (!((xi <= u || i == a) && (u < xi1 || i == b - 1))) {
                    int lo_s = a, hi_s = b - 1;             // largest i in [a, b-1] with x_i <= u (a if none)
                    while (lo_s < hi_s) {
                        const int mid = (lo_s + hi_s + 1) >> 1;
                        if (__dmul_rn(k, __dsub_rn(STT(mid), al)) <= u) lo_s = mid; else hi_s = mid - 1;
                    }
                    i = lo_s;
                    xi = __dmul_rn(k, __dsub_rn(STT(i), al));
                    xi1 = __dmul_rn(k, __dsub_rn(STT(i + 1), al));
                }
                const double ht = STT(i + 1) - STT(i);
                const double hx = xi1 - xi;
                const double inv_h = 1.0 / hx;
                const double A = (xi1 - u) * inv_h;
                const double B = (u - xi) * inv_h;
                const double h26 = ht * ht * (1.0 / 6.0);
                const double ca = (A * A * A - A) * h26;
                const double cb = (B * B * B - B) * h26;
                const double2 yi = *reinterpret_cast<const double2*>(&SY(i, 2 * c));
                const double2 yi1 = *reinterpret_cast<const double2*>(&SY(i + 1, 2 * c));
                const double2 Mi = *reinterpret_cast<const double2*>(&SD(i, 2 * c));
                const double2 Mi1 = *reinterpret_cast<const double2*>(&SD(i + 1, 2 * c));
                double2 r;
                r.x = A * yi.x + B * yi1.x + (ca * Mi.x + cb * Mi1.x);
                r.y = A * yi.y + B * yi1.y + (ca * Mi.y + cb * Mi1.y);
                const int64_t o = (tshift > 0) ? (int64_t)(j >> tshift) * tileGT + ((int64_t)g << tshift) + (j & tmask)
                                               : (int64_t)j * G + g;
                o2[o] = r;
            }
        }
    } else {
        // one thread per (knot, real column): 128-byte row segments of the time-major output
        const int rstep = nthr / ST_COLS;
        const bool colok = col0 + cc < G2;
        for (int i = a + sidx; i < b; i += rstep) {
            const double ht = STT(i + 1) - STT(i);
            const double yi = SY(i, cc), yi1 = SY(i + 1, cc);
            const double Mi = SD(i, cc), Mi1 = SD(i + 1, cc);
            if (!colok) continue;
            if (MODE == 1) {
                const double dl = (yi1 - yi) / ht, h6 = ht * (1.0 / 6.0);
                out[(size_t)i * G2 + col0 + cc] = dl - h6 * (2.0 * Mi + Mi1);
                if (i == N - 2) out[(size_t)(N - 1) * G2 + col0 + cc] = dl + h6 * (Mi + 2.0 * Mi1);
            } else if (MODE == 2) {
                out[(size_t)i * G2 + col0 + cc] = Mi;
                if (i == N - 2) out[(size_t)(N - 1) * G2 + col0 + cc] = Mi1;
            } else if (MODE == 3) {
                out[(size_t)(i + 1) * G2 + col0 + cc] = 0.5 * ht * (yi + yi1) - (ht * ht * ht) * (1.0 / 24.0) * (Mi + Mi1);
                if (i == 0) out[col0 + cc] = 0.0;
            } else {
                // second antiderivative over the interval: II_{i+1} - II_i = h I_i + int int of the cubic piece
                const double c1 = (yi1 - yi) / ht - ht * (1.0 / 6.0) * (2.0 * Mi + Mi1);
                const double h2 = ht * ht;
                const double J = h2 * (0.5 * yi + ht * (1.0 / 6.0) * c1 + h2 * ((1.0 / 24.0) * Mi + (1.0 / 120.0) * (Mi1 - Mi)));
                out[(size_t)(i + 1) * G2 + col0 + cc] = J + ht * up[(size_t)i * G2 + col0 + cc];
                if (i == 0) out[col0 + cc] = 0.0;
            }
        }
    }
#undef SY
#undef SD
#undef STAB
#undef STT
}

// Inclusive scan down the columns of x[N, C] (in place), one CTA per 32 real columns, rows in slabs: used to turn the
// per-interval integrals into the antiderivative (waveform_base.py:697-703).  C is small (hundreds), N long: each
// thread owns one column and a contiguous slab of rows; slab totals are combined through shared memory.
__global__ void __launch_bounds__(1024)
column_scan_kernel(double* __restrict__ x, int64_t N, int C) {
    __shared__ double s_tot[32][33];
    const int col = blockIdx.x * 32 + threadIdx.x;
    const int slab = threadIdx.y;                   // 0..31
    const int64_t per = (N + 31) / 32;
    const int64_t r0 = slab * per, r1 = (r0 + per < N) ? r0 + per : N;
    double acc = 0.0;
    if (col < C)
        for (int64_t r = r0; r < r1; ++r) acc += x[r * C + col];
    s_tot[slab][threadIdx.x] = acc;
    __syncthreads();
    double base = 0.0;
    for (int s = 0; s < slab; ++s) base += s_tot[s][threadIdx.x];
    if (col < C) {
        double run = base;
        for (int64_t r = r0; r < r1; ++r) {
            run += x[r * C + col];
            x[r * C + col] = run;
        }
    }
}

static size_t tile_smem_bytes(int body, int halo) {
    const size_t ymax = (size_t)body + 2 * halo + 3, dmax = (size_t)body + halo + 4;
    return (ymax + dmax) * ST_PITCH * sizeof(double) + ymax * (sizeof(double4) + sizeof(double));
}

template <int MODE>
static int launch_tile(const double* t, int64_t n_times, const double* F, int G, const double* kconf, const double* alpha,
                       const double* tab, const double* uprm, int64_t n_out, double* out, int tshift, int halo, int body,
                       void* stream, const char* name) {
    SCRIB200_REQUIRE(n_times < (int64_t)2147483000 && n_out < (int64_t)2147483000, "%s: series longer than 2^31 samples", name);
    if (halo <= 0) halo = 32;
    if (body <= 0) body = (halo <= 32) ? 256 : (halo <= 64 ? 192 : 128);
    const size_t smem = tile_smem_bytes(body, halo);
    SCRIB200_REQUIRE(halo <= 256 && body >= 16 && smem <= 220 * 1024, "%s: body=%d halo=%d does not fit shared memory", name, body, halo);
    const int64_t ntiles = (n_times - 1 + body - 1) / body;
    SCRIB200_REQUIRE(ntiles <= 65535, "%s: too many time tiles (%lld); raise `body`", name, (long long)ntiles);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(spline_tile_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((2 * G + ST_COLS - 1) / ST_COLS, (unsigned)ntiles);
    int threads = 256;
    if (const char* env = getenv("SCRIB200_SPLINE_THREADS")) {   // tuning knob (multiple of 32, 64..256)
        const int v = atoi(env);
        if (v >= 64 && v <= 256 && v % 32 == 0) threads = v;
    }
    spline_tile_kernel<MODE><<<grid, threads, smem, (cudaStream_t)stream>>>(
        t, (int)n_times, F, G, kconf, alpha, reinterpret_cast<const double4*>(tab), uprm, (int)n_out, out, tshift, body, halo);
    SCRIB200_CHECK_LAUNCH(name);
    return SCRIB200_OK;
}

}  // namespace scrib200

extern "C" int scrib200_spline_prepare(const double* t, int64_t n_times, double inv_gamma, double time_translation,
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
    spline_factor_kernel<<<blocks, 256, 0, st>>>(t, (int)n_times, reinterpret_cast<double4*>(tab), inv_gamma,
                                                 time_translation, kconf, alpha, G, uprm, info);
    SCRIB200_CHECK_LAUNCH("spline_prepare(factor)");
    spline_decay_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const double4*>(tab), (int)n_times, info);
    SCRIB200_CHECK_LAUNCH("spline_prepare(decay)");
    return SCRIB200_OK;
}

extern "C" int scrib200_spline_remap(const double* t, int64_t n_times, const double* F, int G, const double* kconf,
                                     const double* alpha, const double* tab, const double* uprm, int64_t n_out,
                                     double* out, int tile, int halo, int body, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(t && F && kconf && alpha && tab && uprm && out, "spline_remap: null pointer");
    SCRIB200_REQUIRE(n_times >= 4, "spline_remap: a cubic interpolating spline needs at least 4 knots; got %lld",
                     (long long)n_times);
    int tshift = 0;
    while ((1 << tshift) < tile) ++tshift;
    SCRIB200_REQUIRE(G > 0 && (tile == 0 || (tile >= 2 && (1 << tshift) == tile)),
                     "spline_remap: G=%d, tile=%d must be 0 (time-major output) or a power of two >= 2", G, tile);
    SCRIB200_REQUIRE(aligned16(F) && aligned16(out) && aligned16(tab), "spline_remap: pointers must be 16-byte aligned");
    if (n_out <= 0) return SCRIB200_OK;
    return launch_tile<0>(t, n_times, F, G, kconf, alpha, tab, uprm, n_out, out, tile ? tshift : 0, halo, body, stream,
                          "spline_remap");
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
        return launch_tile<1>(t, n_times, data, ncol, nullptr, nullptr, tab, nullptr, 0, out, 0, halo, body, stream, "spline_calculus");
    if (order == 2)
        return launch_tile<2>(t, n_times, data, ncol, nullptr, nullptr, tab, nullptr, 0, out, 0, halo, body, stream, "spline_calculus");
    const int C = 2 * ncol;
    double* first = (order == -1) ? out : aux;
    int rc = launch_tile<3>(t, n_times, data, ncol, nullptr, nullptr, tab, nullptr, 0, first, 0, halo, body, stream, "spline_calculus");
    if (rc != SCRIB200_OK) return rc;
    column_scan_kernel<<<(C + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream>>>(first, n_times, C);
    SCRIB200_CHECK_LAUNCH("spline_calculus(scan)");
    if (order == -1) return SCRIB200_OK;
    rc = launch_tile<4>(t, n_times, data, ncol, nullptr, nullptr, tab, first, 0, out, 0, halo, body, stream, "spline_calculus");
    if (rc != SCRIB200_OK) return rc;
    column_scan_kernel<<<(C + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream>>>(out, n_times, C);
    SCRIB200_CHECK_LAUNCH("spline_calculus(scan)");
    return SCRIB200_OK;
}
