// K5/K6/K7 and friends: per-time-step reductions over the modes (mode_calculations.py, flux.py).
//
//   norm                 scri/waveform_base.py:19-35          sum_lm |a|^2
//   ll_ldt               scri/mode_calculations.py:14-43,209-295   <LL> (3x3 real sym.) and <L d/dt> (3-vector)
//   l_vector             scri/mode_calculations.py:60-89      <L> (complex 3-vector)
//   sym3_dominant_eig    np.linalg.eigh + [:, :, 2]           (mode_calculations.py:389-395)
//   eig_sign_*           scri/mode_calculations.py:316-363    sequential sign-continuity pass as a 2-state scan
//   solve3               np.linalg.solve (mode_calculations.py:424)
//   sparse_expectation   scri/flux.py:40-78                   sum_e conj(a[row_e]) b[col_e] val_e
//
// All are HBM-bound: one warp per time step, lanes striding over modes (coalesced 16-byte loads of the
// [time, mode] row), warp-shuffle reduction, a few bytes out per step.  The reference loops mode-outer /
// time-inner with stride-n access; here each row is read exactly once.
#include <algorithm>

#include "common.cuh"

namespace scrib200 {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------ norm
// eight lanes per time step, four steps per warp: at ell <= 8 (77 modes) a whole warp per step leaves most lanes with two
// loads and a five-level reduction per step
__global__ void norm_kernel(const double2* __restrict__ data, int64_t n_times, int n, double* __restrict__ out) {
    const int lane = threadIdx.x & 7;
    int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool live = t < n_times;
    if (!live) t = n_times - 1;                          // keeps the whole warp in the shuffles
    const double2* row = data + t * n;
    double acc = 0.0;
#pragma unroll 4
    for (int i = lane; i < n; i += 8) {
        const double2 a = row[i];
        acc = fma(a.x, a.x, acc);
        acc = fma(a.y, a.y, acc);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && live) out[t] = acc;
}

// ------------------------------------------------------------------------------------------ <LL>, <L dt>
// coef[i][0..4] = { ladder(l,m) [0 if m+1>l], ladder(l,-m) [0 if m-1<-l], ladder(l,m+1)ladder(l,m) [0 if m+2>l],
//                   ladder(l,-(m-1))ladder(l,-m) [0 if m-2<-l], m }
constexpr int NCOEF = 5;

// Eight lanes per time step, four time steps per warp: the nine sums of a step are reduced over 8 lanes (3 butterfly levels
// instead of 5) for four steps at once - with a whole warp per step the reductions, not the arithmetic, were most of the time
// (9 modes per lane at ell <= 16, 3 at ell <= 8).
constexpr int LL_GROUP = 8;

__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = LL_GROUP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool WITH_LDT>
__global__ void ll_ldt_kernel(const double2* __restrict__ data, const double2* __restrict__ datadot, int64_t n_times,
                              int n, const double* __restrict__ coef, double* __restrict__ LL, double* __restrict__ Ldt) {
    const int lane = threadIdx.x & (LL_GROUP - 1);
    int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LL_GROUP;
    const bool live = t < n_times;
    if (!live) t = n_times - 1;                          // keeps the whole warp in the shuffles
    const double2* row = data + t * n;
    const double2* rowd = WITH_LDT ? datadot + t * n : nullptr;
    double sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0, lx = 0, ly = 0, lz = 0;
    for (int i = lane; i < n; i += LL_GROUP) {
        const double* c = coef + i * NCOEF;
        const double cp = c[0], cm = c[1], cpp = c[2], cmm = c[3], M = c[4];
        const double2 a = row[i];
        const double2 ap1 = cconj(row[i + 1 < n ? i + 1 : i]);
        const double2 am1 = cconj(row[i >= 1 ? i - 1 : i]);
        const double2 ap2 = cconj(row[i + 2 < n ? i + 2 : i]);
        const double2 am2 = cconj(row[i >= 2 ? i - 2 : i]);
        const double aa = a.x * a.x + a.y * a.y;
        // products in the (+,-,z) basis, as in the reference
        const double2 LpLp = cscale(cpp, cmul(ap2, a));
        const double2 LmLm = cscale(cmm, cmul(am2, a));
        const double LpLm = (cm != 0.0) ? aa * (cm * cm) : 0.0;   // ladder(l,m-1)*ladder(l,-m) = ladder(l,-m)^2
        const double LmLp = (cp != 0.0) ? aa * (cp * cp) : 0.0;   // ladder(l,-(m+1))*ladder(l,m) = ladder(l,m)^2
        const double2 p1 = cmul(ap1, a), m1 = cmul(am1, a);
        const double2 LpLz = cscale(cp * M, p1);
        const double2 LzLp = cscale((M + 1.0) * cp, p1);
        const double2 LmLz = cscale(cm * M, m1);
        const double2 LzLm = cscale((M - 1.0) * cm, m1);
        const double LzLz = aa * (M * M);
        // real parts of the symmetrised (x,y,z) products
        sxx += 0.25 * (LpLp.x + LmLm.x + LmLp + LpLm);
        // LxLy + LyLx = -0.25j*(2 LpLp - 2 LmLm)  -> real part = 0.5*(Im LpLp - Im LmLm); halved by the symmetrisation
        sxy += 0.25 * (LpLp.y - LmLm.y);
        sxz += 0.25 * ((LpLz.x + LmLz.x) + (LzLp.x + LzLm.x));
        syy += -0.25 * (LpLp.x - LmLp - LpLm + LmLm.x);
        // LyLz + LzLy = -0.5j*((LpLz - LmLz) + (LzLp - LzLm)) -> real part = 0.5*Im(...); halved
        syz += 0.25 * ((LpLz.y - LmLz.y) + (LzLp.y - LzLm.y));
        szz += LzLz;
        if (WITH_LDT) {
            const double2 ad = rowd[i];
            const double2 Lp = cscale(cp, cmul(ap1, ad));
            const double2 Lm = cscale(cm, cmul(am1, ad));
            const double2 Lz = cscale(M, cmul(cconj(a), ad));
            lx += 0.5 * (Lp.y + Lm.y);
            ly += -0.5 * (Lp.x - Lm.x);
            lz += Lz.y;
        }
    }
    sxx = group_sum(sxx); sxy = group_sum(sxy); sxz = group_sum(sxz);
    syy = group_sum(syy); syz = group_sum(syz); szz = group_sum(szz);
    if (WITH_LDT) { lx = group_sum(lx); ly = group_sum(ly); lz = group_sum(lz); }
    if (lane == 0 && live) {
        double* o = LL + t * 9;
        o[0] = sxx; o[1] = sxy; o[2] = sxz;
        o[3] = sxy; o[4] = syy; o[5] = syz;
        o[6] = sxz; o[7] = syz; o[8] = szz;
        if (WITH_LDT) {
            double* v = Ldt + t * 3;
            v[0] = lx; v[1] = ly; v[2] = lz;
        }
    }
}

// <LL>^{ab} between two waveforms (mode_calculations.py:106-206), complex and NOT symmetrised, reproduced literally -
// including the reference's accumulation of the (y,z) element into (y,y) (LL[1,1] receives two terms, LL[1,2] none).
__global__ void ll_comparison_kernel(const double2* __restrict__ d1, const double2* __restrict__ d2, int64_t n_times, int n,
                                     const double* __restrict__ coef, double2* __restrict__ LL) {
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_times) return;
    const double2* r1 = d1 + t * n;
    const double2* r2 = d2 + t * n;
    double2 s[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) s[q] = make_double2(0.0, 0.0);
    for (int i = lane; i < n; i += 32) {
        const double* c = coef + i * NCOEF;
        const double cp = c[0], cm = c[1], cpp = c[2], cmm = c[3], M = c[4];
        const double2 g = r2[i];
        const double2 f0 = cmul(cconj(r1[i]), g);
        const double2 fp1 = cmul(cconj(r1[i + 1 < n ? i + 1 : i]), g), fm1 = cmul(cconj(r1[i >= 1 ? i - 1 : i]), g);
        const double2 fp2 = cmul(cconj(r1[i + 2 < n ? i + 2 : i]), g), fm2 = cmul(cconj(r1[i >= 2 ? i - 2 : i]), g);
        const double2 LpLp = cscale(cpp, fp2), LmLm = cscale(cmm, fm2);
        const double2 LpLm = cscale(cm * cm, f0), LmLp = cscale(cp * cp, f0);
        const double2 LpLz = cscale(cp * M, fp1), LzLp = cscale((M + 1.0) * cp, fp1);
        const double2 LmLz = cscale(cm * M, fm1), LzLm = cscale((M - 1.0) * cm, fm1);
        const double2 LzLz = cscale(M * M, f0);
        auto add = [](double2& acc, double re, double im) { acc.x += re; acc.y += im; };
        // 0.25 (LpLp + LmLm + LmLp + LpLm)
        add(s[0], 0.25 * (LpLp.x + LmLm.x + LmLp.x + LpLm.x), 0.25 * (LpLp.y + LmLm.y + LmLp.y + LpLm.y));
        {   // -0.25j (LpLp - LmLm + LmLp - LpLm):  -j (x + i y) = y - i x
            const double x = LpLp.x - LmLm.x + LmLp.x - LpLm.x, y = LpLp.y - LmLm.y + LmLp.y - LpLm.y;
            add(s[1], 0.25 * y, -0.25 * x);
        }
        add(s[2], 0.5 * (LpLz.x + LmLz.x), 0.5 * (LpLz.y + LmLz.y));
        {   // -0.25j (LpLp - LmLp + LpLm - LmLm)
            const double x = LpLp.x - LmLp.x + LpLm.x - LmLm.x, y = LpLp.y - LmLp.y + LpLm.y - LmLm.y;
            add(s[3], 0.25 * y, -0.25 * x);
        }
        add(s[4], -0.25 * (LpLp.x - LmLp.x - LpLm.x + LmLm.x), -0.25 * (LpLp.y - LmLp.y - LpLm.y + LmLm.y));
        {   // the reference adds -0.5j (LpLz - LmLz) to element (1,1) as well
            const double x = LpLz.x - LmLz.x, y = LpLz.y - LmLz.y;
            add(s[4], 0.5 * y, -0.5 * x);
        }
        add(s[6], 0.5 * (LzLp.x + LzLm.x), 0.5 * (LzLp.y + LzLm.y));
        {   // -0.5j (LzLp - LzLm)
            const double x = LzLp.x - LzLm.x, y = LzLp.y - LzLm.y;
            add(s[7], 0.5 * y, -0.5 * x);
        }
        add(s[8], LzLz.x, LzLz.y);
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        s[q].x = warp_sum(s[q].x);
        s[q].y = warp_sum(s[q].y);
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 9; ++q) LL[t * 9 + q] = s[q];
    }
}

// <L>^a = sum conj(a1[m']) <L_a> a2[m]  (complex 3-vector)
__global__ void l_vector_kernel(const double2* __restrict__ d1, const double2* __restrict__ d2, int64_t n_times, int n,
                                const double* __restrict__ coef, double2* __restrict__ Lvec) {
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_times) return;
    const double2* r1 = d1 + t * n;
    const double2* r2 = d2 + t * n;
    double2 sx = make_double2(0, 0), sy = sx, sz = sx;
    for (int i = lane; i < n; i += 32) {
        const double* c = coef + i * NCOEF;
        const double2 b = r2[i];
        const double2 Lp = cscale(c[0], cmul(cconj(r1[i + 1 < n ? i + 1 : i]), b));
        const double2 Lm = cscale(c[1], cmul(cconj(r1[i >= 1 ? i - 1 : i]), b));
        const double2 Lz = cscale(c[4], cmul(cconj(r1[i]), b));
        sx.x += 0.5 * (Lp.x + Lm.x); sx.y += 0.5 * (Lp.y + Lm.y);
        // -0.5j * (Lp - Lm)
        sy.x += 0.5 * (Lp.y - Lm.y); sy.y += -0.5 * (Lp.x - Lm.x);
        sz.x += Lz.x; sz.y += Lz.y;
    }
    sx.x = warp_sum(sx.x); sx.y = warp_sum(sx.y);
    sy.x = warp_sum(sy.x); sy.y = warp_sum(sy.y);
    sz.x = warp_sum(sz.x); sz.y = warp_sum(sz.y);
    if (lane == 0) {
        Lvec[t * 3 + 0] = sx; Lvec[t * 3 + 1] = sy; Lvec[t * 3 + 2] = sz;
    }
}

// ------------------------------------------------------------------------------------------ 3x3 symmetric eigenvector
// Cyclic Jacobi on the 3x3 symmetric matrix; returns the unit eigenvector of the largest eigenvalue.
__device__ void sym3_dominant(const double* A, double* v) {
    double a[3][3] = {{A[0], A[1], A[2]}, {A[3], A[4], A[5]}, {A[6], A[7], A[8]}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-300 || off <= 1e-18 * diag) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0;
            const int q = (pq == 0) ? 1 : 2;
            const double apq = a[p][q];
            if (apq == 0.0) continue;
            const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
            const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
            const int r = 3 - p - q;
            const double app = a[p][p], aqq = a[q][q];
            a[p][p] = app - tt * apq;
            a[q][q] = aqq + tt * apq;
            a[p][q] = a[q][p] = 0.0;
            const double arp = a[r][p], arq = a[r][q];
            a[r][p] = a[p][r] = c * arp - s * arq;
            a[r][q] = a[q][r] = s * arp + c * arq;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double vkp = V[k][p], vkq = V[k][q];
                V[k][p] = c * vkp - s * vkq;
                V[k][q] = s * vkp + c * vkq;
            }
        }
    }
    int best = 0;
    if (a[1][1] > a[best][best]) best = 1;
    if (a[2][2] > a[best][best]) best = 2;
    double x = V[0][best], y = V[1][best], z = V[2][best];
    const double nrm = sqrt(x * x + y * y + z * z);
    v[0] = x / nrm; v[1] = y / nrm; v[2] = z / nrm;
}

__global__ void sym3_dominant_eig_kernel(const double* __restrict__ LL, int64_t n_times, double* __restrict__ dpa) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_times) return;
    double A[9], v[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k] = LL[t * 9 + k];
    sym3_dominant(A, v);
    dpa[t * 3 + 0] = v[0]; dpa[t * 3 + 1] = v[1]; dpa[t * 3 + 2] = v[2];
}

// ---- sign continuity (mode_calculations.py:316-363) as a scan over 2-state maps.
// For step i with predecessor j (= i-1 going forward from i_index, i+1 going backward) the reference flips
// raw v_i iff |v_i - s_j v_j|^2 > |v_i|^2, s_j the sign already applied to v_j.  map bit0 = flip if s_j=+1,
// bit1 = flip if s_j=-1.
__global__ void eig_sign_map_kernel(const double* __restrict__ dpa, int64_t n, int64_t i_index, unsigned char* __restrict__ maps) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == i_index) { maps[i] = 0; return; }
    const int64_t j = (i > i_index) ? i - 1 : i + 1;
    const double x = dpa[i * 3], y = dpa[i * 3 + 1], z = dpa[i * 3 + 2];
    const double px = dpa[j * 3], py = dpa[j * 3 + 1], pz = dpa[j * 3 + 2];
    const double Norm = x * x + y * y + z * z;
    const double dp = (x - px) * (x - px) + (y - py) * (y - py) + (z - pz) * (z - pz);
    const double dm = (x + px) * (x + px) + (y + py) * (y + py) + (z + pz) * (z + pz);
    maps[i] = (unsigned char)((dp > Norm ? 1 : 0) | (dm > Norm ? 2 : 0));
}

// The walk s_i = f_i(s_{i-1}) over the 2-state maps is a scan over function composition, done in three parallel
// phases per direction (forward from i_index+1, backward from i_index-1): (A) every thread composes a block of
// SIGN_BLK consecutive maps into one map; (B) one CTA stages the block maps in shared memory and a single thread per
// direction walks them (n / SIGN_BLK cheap steps) recording the state entering every block; (C) every thread re-walks
// its block from that state and writes the signs.
constexpr int SIGN_BLK = 256;

__device__ __forceinline__ int64_t sign_len(int dir, int64_t n, int64_t i_index) { return dir == 0 ? n - 1 - i_index : i_index; }
__device__ __forceinline__ int64_t sign_pos(int dir, int64_t i_index, int64_t p) { return dir == 0 ? i_index + 1 + p : i_index - 1 - p; }
// state: 0 = sign +1, 1 = sign -1;  next state under map m
__device__ __forceinline__ int sign_next(unsigned char m, int s) { return s == 0 ? (m & 1) : ((m >> 1) & 1); }

__global__ void eig_sign_blocks_kernel(const unsigned char* __restrict__ maps, int64_t n, int64_t i_index, int64_t nblk_dir,
                                       unsigned char* __restrict__ blkmap) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 2 * nblk_dir) return;
    const int dir = id >= nblk_dir;
    const int64_t bidx = id - dir * nblk_dir;
    const int64_t len = sign_len(dir, n, i_index);
    int a = 0, b = 1;   // images of +1 and -1
    const int64_t p1 = (bidx + 1) * SIGN_BLK < len ? (bidx + 1) * SIGN_BLK : len;
    for (int64_t p = bidx * SIGN_BLK; p < p1; ++p) {
        const unsigned char m = maps[sign_pos(dir, i_index, p)];
        a = sign_next(m, a);
        b = sign_next(m, b);
    }
    blkmap[id] = (unsigned char)(a | (b << 1));
}

__global__ void eig_sign_walk_kernel(const double* __restrict__ dpa, int64_t i_index, const double* __restrict__ rough,
                                     int64_t nblk_dir, const unsigned char* __restrict__ blkmap,
                                     unsigned char* __restrict__ blkstate, signed char* __restrict__ signs) {
    extern __shared__ unsigned char s_blk[];
    const double d0 = rough[0] * dpa[i_index * 3] + rough[1] * dpa[i_index * 3 + 1] + rough[2] * dpa[i_index * 3 + 2];
    const int s0 = (d0 < 0.0) ? 1 : 0;
    if (threadIdx.x == 0) signs[i_index] = s0 ? -1 : 1;
    // walk in slabs that fit shared memory
    const int64_t SLAB = 16384;
    int state[2] = {s0, s0};
    for (int64_t base = 0; base < nblk_dir; base += SLAB) {
        const int64_t cnt = nblk_dir - base < SLAB ? nblk_dir - base : SLAB;
        __syncthreads();
        for (int64_t e = threadIdx.x; e < 2 * cnt; e += blockDim.x) {
            const int dir = e >= cnt;
            s_blk[e] = blkmap[dir * nblk_dir + base + (e - dir * cnt)];
        }
        __syncthreads();
        if (threadIdx.x < 2) {
            const int dir = threadIdx.x;
            int s = state[dir];
            for (int64_t q = 0; q < cnt; ++q) {
                const unsigned char m = s_blk[dir * cnt + q];
                s_blk[dir * cnt + q] = (unsigned char)s;          // the state entering block q
                s = sign_next(m, s);
            }
            state[dir] = s;
        }
        __syncthreads();
        for (int64_t e = threadIdx.x; e < 2 * cnt; e += blockDim.x) {
            const int dir = e >= cnt;
            blkstate[dir * nblk_dir + base + (e - dir * cnt)] = s_blk[e] & 1;
        }
        // thread 0 / 1 carry `state` themselves; the others never use it
    }
}

__global__ void eig_sign_fill_kernel(const unsigned char* __restrict__ maps, int64_t n, int64_t i_index, int64_t nblk_dir,
                                     const unsigned char* __restrict__ blkstate, signed char* __restrict__ signs) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 2 * nblk_dir) return;
    const int dir = id >= nblk_dir;
    const int64_t bidx = id - dir * nblk_dir;
    const int64_t len = sign_len(dir, n, i_index);
    int s = blkstate[id];
    const int64_t p1 = (bidx + 1) * SIGN_BLK < len ? (bidx + 1) * SIGN_BLK : len;
    for (int64_t p = bidx * SIGN_BLK; p < p1; ++p) {
        const int64_t i = sign_pos(dir, i_index, p);
        s = sign_next(maps[i], s);
        signs[i] = s ? -1 : 1;
    }
}

__global__ void eig_sign_finish_kernel(double* __restrict__ dpa, int64_t n, const signed char* __restrict__ signs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = dpa[i * 3], y = dpa[i * 3 + 1], z = dpa[i * 3 + 2];
    const double s = (double)signs[i];
    x *= s; y *= s; z *= s;
    const double nrm = sqrt(x * x + y * y + z * z);
    if (nrm != 0.0 && nrm != 1.0) { x /= nrm; y /= nrm; z /= nrm; }
    dpa[i * 3] = x; dpa[i * 3 + 1] = y; dpa[i * 3 + 2] = z;
}

// ------------------------------------------------------------------------------------------ 3x3 solve (LU, partial pivoting as gesv)
__global__ void solve3_kernel(const double* __restrict__ A9, const double* __restrict__ b3, int64_t n, double scale,
                              double* __restrict__ x3) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double a[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) a[r][c] = A9[t * 9 + r * 3 + c];
        a[r][3] = b3[t * 3 + r];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        for (int r = k + 1; r < 3; ++r)
            if (fabs(a[r][k]) > fabs(a[piv][k])) piv = r;
        if (piv != k) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { const double tmp = a[k][c]; a[k][c] = a[piv][c]; a[piv][c] = tmp; }
        }
        for (int r = k + 1; r < 3; ++r) {
            const double f = a[r][k] / a[k][k];
#pragma unroll
            for (int c = k; c < 4; ++c) a[r][c] -= f * a[k][c];
        }
    }
    double x[3];
    x[2] = a[2][3] / a[2][2];
    x[1] = (a[1][3] - a[1][2] * x[2]) / a[1][1];
    x[0] = (a[0][3] - a[0][1] * x[1] - a[0][2] * x[2]) / a[0][0];
    x3[t * 3] = scale * x[0]; x3[t * 3 + 1] = scale * x[1]; x3[t * 3 + 2] = scale * x[2];
}

// ------------------------------------------------------------------------------------------ sparse <a|M|b>
// K matrices share one pass over the rows: elements of matrix k are [seg[k], seg[k+1]).  A warp owns ST consecutive
// time steps, so every (row, col, value) triple fetched serves ST products: the kernel is bound by load issue, not
// by HBM, and the index / value loads were 3 of the 5 loads per element.
constexpr int SPARSE_ST = 1;   // measured: 2 and 4 steps per warp are slower (fewer warps in flight outweigh the saved index loads)

__global__ void __launch_bounds__(128)
sparse_expectation_kernel(const double2* __restrict__ a, const double2* __restrict__ b, int64_t n_times,
                          int n, const int* __restrict__ rows, const int* __restrict__ cols,
                          const double2* __restrict__ vals, const int* __restrict__ seg, int K,
                          double2* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t t0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * SPARSE_ST;
    if (t0 >= n_times) return;
    const int nst = (n_times - t0 < SPARSE_ST) ? (int)(n_times - t0) : SPARSE_ST;
    const double2* ra = a + t0 * n;
    const double2* rb = b + t0 * n;
    for (int k = 0; k < K; ++k) {
        double2 acc[SPARSE_ST];
#pragma unroll
        for (int s = 0; s < SPARSE_ST; ++s) acc[s] = make_double2(0.0, 0.0);
        if (nst == SPARSE_ST) {
            for (int e = seg[k] + lane; e < seg[k + 1]; e += 32) {
                const int r = rows[e], c = cols[e];
                const double2 v = vals[e];
#pragma unroll
                for (int s = 0; s < SPARSE_ST; ++s) {
                    const double2 p = cmul(cconj(ra[(size_t)s * n + r]), rb[(size_t)s * n + c]);
                    cfma(acc[s], p, v);
                }
            }
        } else {
            for (int e = seg[k] + lane; e < seg[k + 1]; e += 32) {
                const int r = rows[e], c = cols[e];
                const double2 v = vals[e];
                for (int s = 0; s < nst; ++s) {
                    const double2 p = cmul(cconj(ra[(size_t)s * n + r]), rb[(size_t)s * n + c]);
                    cfma(acc[s], p, v);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < SPARSE_ST; ++s) {
            acc[s].x = warp_sum(acc[s].x);
            acc[s].y = warp_sum(acc[s].y);
            if (lane == 0 && s < nst) out[(t0 + s) * K + k] = acc[s];
        }
    }
}

// Column-major ELL form of the same contraction (round 2).  The flux matrices are banded in (l, m): every column has at most
// a handful of entries (3 for the momentum operators, 1 for L_z, ...).  Entry w of column c of matrix k sits at
// [off_k + w n + c]: row index (or -1) and value.  A lane owns columns c = lane, lane + 32, ...: b[c] is loaded once and
// serves every entry of the column in up to four matrices, the index and value loads are coalesced along c, and only a[r]
// is a gather (into the 16 n bytes of the time step, next to its neighbours' rows).  3 loads per non-zero instead of 5
// for the COO kernel, which is bound by load issue.
struct EllLayout {
    int width[32];
    int offset[32];   // in entries
};

template <int KB, bool REALV>
__global__ void __launch_bounds__(128)
sparse_expectation_ell_kernel(const double2* __restrict__ a, const double2* __restrict__ b, int64_t n_times, int n,
                              const int* __restrict__ ell_r, const double* __restrict__ ell_v, EllLayout lay, int k0, int K,
                              double2* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_times) return;
    const double2* ra = a + t * n;
    const double2* rb = b + t * n;
    double2 acc[KB];
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) acc[kk] = make_double2(0.0, 0.0);
    for (int c = lane; c < n; c += 32) {
        const double2 bc = rb[c];
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
            if (k0 + kk < K) {
                const int W = lay.width[k0 + kk];
                const int* rr = ell_r + lay.offset[k0 + kk] + c;
                const double* vv = ell_v + (size_t)(REALV ? 1 : 2) * (lay.offset[k0 + kk] + c);
                for (int w = 0; w < W; ++w) {
                    const int r = rr[(size_t)w * n];
                    if (r >= 0) {
                        const double2 p = cmul(cconj(ra[r]), bc);
                        if (REALV) {
                            const double v = vv[(size_t)w * n];
                            acc[kk].x = fma(p.x, v, acc[kk].x);
                            acc[kk].y = fma(p.y, v, acc[kk].y);
                        } else {
                            const double2 v = *reinterpret_cast<const double2*>(vv + (size_t)2 * w * n);
                            cfma(acc[kk], p, v);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
        acc[kk].x = warp_sum(acc[kk].x);
        acc[kk].y = warp_sum(acc[kk].y);
        if (lane == 0 && k0 + kk < K) out[t * K + k0 + kk] = acc[kk];
    }
}

// Time-lane form of the same contraction.  Lanes run along TIME: a CTA owns 32 time steps, stages the mode rows it needs
// transposed in shared memory ([mode][time], pitch 33: conflict-free both ways) and walks the columns of the matrices with
// everything that steers the walk - row index, value, emptiness of a slot - uniform across the warp (one broadcast
// 16-byte load per non-zero instead of per-lane index / value loads and a gather).  The non-zeros of column c of matrix k
// are first folded into q = sum_w v_w conj(a[r_w]) (2 FMAs each), then acc_k += q b[c] (4 FMAs per column and matrix): for the
// momentum operators (3 entries per column) that is 3.3 FP64 instructions per non-zero where the warp-per-time-step
// kernels issue ~40 instructions of all kinds.  The columns are cut into blocks whose rows of `a` fit shared memory (the
// flux matrices are banded in (l, m): scri/flux.py:213-298), computed on the host with the tables
// (scri_b200/ops.py:_time_tables); the warps of a CTA split the columns of a block and their sums are added at the end.
constexpr int XT_T = 32;
constexpr int XT_PITCH = 33;
constexpr int XT_WARPS = 8;
constexpr int XT_MAXBLK = 32;
constexpr int XT_MAXW = 4;     // entries per column and matrix the unrolled walk covers
struct XtPlan {
    int n_blocks;
    int c0[XT_MAXBLK], c1[XT_MAXBLK], rlo[XT_MAXBLK], rhi[XT_MAXBLK];
    int width[4];              // slots of the (up to) four matrices of this launch
    int slot0;                 // first slot of matrix k0 within a column
    int slots;                 // slots per column (all matrices)
    int max_rows, max_cols;    // shared-memory tile heights
};

__device__ __forceinline__ void xt_cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

// stage rows [m0, m1) of x for the CTA's 32 time steps, transposed ([mode][time]): a thread keeps its mode and issues the
// 32 asynchronous 16-byte copies of its time steps back to back (coalesced along the modes in global memory, conflict-free
// in shared memory; time steps past the end of the series are zero-filled) - one memory latency per block, no registers
__device__ __forceinline__ void xt_stage(double2* __restrict__ dst, const double2* __restrict__ x, int64_t t0, int64_t n_times, int n,
                                         int m0, int m1, int tid, int nthreads) {
    const int rows = (n_times - t0 < XT_T) ? (int)(n_times - t0) : XT_T;
    for (int lm = tid; lm < m1 - m0; lm += nthreads) {
        const double2* src = x + t0 * n + m0 + lm;
        double2* d = dst + lm * XT_PITCH;
#pragma unroll 8
        for (int tt = 0; tt < XT_T; ++tt) xt_cp_async16(d + tt, tt < rows ? src + (size_t)tt * n : x, tt < rows ? 16 : 0);
    }
}

__device__ __forceinline__ double2 xt_lds(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// CPLX: 32-byte entries {int row, int 0, double re, double im, double 0}; else 16-byte {int row, int 0, double value}.
// Every matrix has WMAX slots per column; an unused slot names a row the column touches anyway, with value 0 (a column
// without any entry has row -1 in all its slots and is skipped), so the walk is free of tests.  Shared memory is
// addressed with 32-bit shared-space addresses (one IMAD per operand).
template <int WMAX, bool CPLX, bool SAME>
__global__ void __launch_bounds__(32 * XT_WARPS, 2)
sparse_expectation_time_kernel(const double2* __restrict__ a, const double2* __restrict__ b, int64_t n_times, int n,
                               const double2* __restrict__ tab, const __grid_constant__ XtPlan plan, int k0, int K, double2* __restrict__ out) {
    constexpr int EW = CPLX ? 2 : 1;                                 // double2 words per entry
    constexpr int KB = 4;
    extern __shared__ double2 smx[];
    double2* sA = smx;                                               // [max_rows][33]
    double2* sB = SAME ? smx : smx + (size_t)plan.max_rows * XT_PITCH;   // [max_cols][33]
    double2* sT = smx + (size_t)(plan.max_rows + (SAME ? 0 : plan.max_cols)) * XT_PITCH;   // [max_cols][slots] entries of the block
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * XT_T;
    const int slots = plan.slots;
    const int nk = K - k0 < KB ? K - k0 : KB;
    double2 acc[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) acc[k] = make_double2(0.0, 0.0);
    const unsigned sA_addr = (unsigned)__cvta_generic_to_shared(sA) + lane * 16;
    const unsigned sB_addr = (unsigned)__cvta_generic_to_shared(sB) + lane * 16;
    const unsigned sT_addr = (unsigned)__cvta_generic_to_shared(sT);
    for (int blk = 0; blk < plan.n_blocks; ++blk) {
        const int c0 = plan.c0[blk], c1 = plan.c1[blk], rlo = plan.rlo[blk], rhi = plan.rhi[blk];
        if (blk > 0) __syncthreads();                                // the previous block's tiles are no longer read
        xt_stage(sA, a, t0, n_times, n, rlo, rhi, tid, 32 * XT_WARPS);
        if (!SAME) xt_stage(sB, b, t0, n_times, n, c0, c1, tid, 32 * XT_WARPS);
        {   // the block's entries: contiguous in the table
            const double2* src = tab + (size_t)c0 * slots * EW;
            const int words = (c1 - c0) * slots * EW;
            for (int i = tid; i < words; i += 32 * XT_WARPS) xt_cp_async16(sT + i, src + i, 16);
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        __syncthreads();
        const unsigned bcol = SAME ? sA_addr + (unsigned)(c0 - rlo) * (XT_PITCH * 16) : sB_addr;
        const unsigned arow = sA_addr - (unsigned)rlo * (XT_PITCH * 16);          // + r * 528
        for (int c = c0 + warp; c < c1; c += XT_WARPS) {
            unsigned e = sT_addr + (unsigned)(((c - c0) * slots + plan.slot0) * (EW * 16));
            const double2 first = xt_lds(e);
            if ((int)(__double_as_longlong(first.x) & 0xffffffffLL) < 0) continue;   // a column no matrix has entries in
            const double2 bc = xt_lds(bcol + (unsigned)(c - c0) * (XT_PITCH * 16));
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                if (k < nk) {
                    // the entries of (column, matrix) first, then the rows they name, then the arithmetic: the loads
                    // overlap instead of queueing behind each other's use
                    double2 ent[WMAX], ent2[WMAX], ar[WMAX];
#pragma unroll
                    for (int w = 0; w < WMAX; ++w) {
                        ent[w] = xt_lds(e + w * (EW * 16));
                        if (CPLX) ent2[w] = xt_lds(e + w * (EW * 16) + 16);
                    }
#pragma unroll
                    for (int w = 0; w < WMAX; ++w) {
                        const unsigned r = (unsigned)(__double_as_longlong(ent[w].x) & 0xffffffffLL);
                        ar[w] = xt_lds(arow + r * (XT_PITCH * 16));
                    }
                    double2 q = make_double2(0.0, 0.0);              // sum_w v_w conj(a[r_w])
#pragma unroll
                    for (int w = 0; w < WMAX; ++w) {
                        if (CPLX) {
                            const double vx = ent[w].y, vy = ent2[w].x;
                            q.x = fma(vx, ar[w].x, fma(vy, ar[w].y, q.x));        // (vx + i vy)(ar.x - i ar.y)
                            q.y = fma(vy, ar[w].x, fma(-vx, ar[w].y, q.y));
                        } else {
                            const double v = ent[w].y;
                            q.x = fma(v, ar[w].x, q.x);
                            q.y = fma(-v, ar[w].y, q.y);
                        }
                    }
                    cfma(acc[k], q, bc);
                    e += WMAX * (EW * 16);
                }
            }
        }
    }
    // add the warps' partial sums: [warp][k][lane] through shared memory (the tiles are free after a barrier)
    __syncthreads();
    double2* red = smx;
#pragma unroll
    for (int k = 0; k < KB; ++k) red[(warp * KB + k) * 32 + lane] = acc[k];
    __syncthreads();
    for (int id = tid; id < KB * 32; id += 32 * XT_WARPS) {
        const int k = id >> 5, tt = id & 31;
        if (k0 + k >= K || t0 + tt >= n_times) continue;
        double2 ssum = make_double2(0.0, 0.0);
#pragma unroll
        for (int w = 0; w < XT_WARPS; ++w) {
            const double2 v = red[(w * KB + k) * 32 + tt];
            ssum.x += v.x;
            ssum.y += v.y;
        }
        out[(t0 + tt) * K + k0 + k] = ssum;
    }
}

static inline unsigned warp_blocks(int64_t n_times) { return (unsigned)((n_times * 32 + 127) / 128); }

}  // namespace scrib200

using namespace scrib200;

extern "C" int scrib200_norm(const double* data, int64_t n_times, int n_modes, double* out, void* stream) {
    SCRIB200_REQUIRE(data && out, "norm: null pointer");
    if (n_times <= 0) return SCRIB200_OK;
    norm_kernel<<<(unsigned)((n_times * 8 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double2*>(data), n_times,
                                                                                         n_modes, out);
    SCRIB200_CHECK_LAUNCH("norm");
    return SCRIB200_OK;
}

extern "C" int scrib200_ll_ldt(const double* data, const double* datadot, int64_t n_times, int n_modes,
                               const double* coef, double* LL, double* Ldt, void* stream) {
    SCRIB200_REQUIRE(data && coef && LL, "ll_ldt: null pointer");
    SCRIB200_REQUIRE((datadot == nullptr) == (Ldt == nullptr), "ll_ldt: datadot and Ldt go together");
    if (n_times <= 0) return SCRIB200_OK;
    const unsigned ll_blocks = (unsigned)((n_times * LL_GROUP + 127) / 128);
    if (datadot)
        ll_ldt_kernel<true><<<ll_blocks, 128, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const double2*>(data), reinterpret_cast<const double2*>(datadot), n_times, n_modes, coef, LL, Ldt);
    else
        ll_ldt_kernel<false><<<ll_blocks, 128, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const double2*>(data), nullptr, n_times, n_modes, coef, LL, nullptr);
    SCRIB200_CHECK_LAUNCH("ll_ldt");
    return SCRIB200_OK;
}

extern "C" int scrib200_l_vector(const double* data1, const double* data2, int64_t n_times, int n_modes,
                                 const double* coef, double* Lvec, void* stream) {
    SCRIB200_REQUIRE(data1 && data2 && coef && Lvec, "l_vector: null pointer");
    if (n_times <= 0) return SCRIB200_OK;
    l_vector_kernel<<<warp_blocks(n_times), 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2*>(data1), reinterpret_cast<const double2*>(data2), n_times, n_modes, coef,
        reinterpret_cast<double2*>(Lvec));
    SCRIB200_CHECK_LAUNCH("l_vector");
    return SCRIB200_OK;
}

extern "C" int scrib200_ll_comparison(const double* data1, const double* data2, int64_t n_times, int n_modes,
                                      const double* coef, double* LL, void* stream) {
    SCRIB200_REQUIRE(data1 && data2 && coef && LL, "ll_comparison: null pointer");
    if (n_times <= 0) return SCRIB200_OK;
    ll_comparison_kernel<<<warp_blocks(n_times), 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2*>(data1), reinterpret_cast<const double2*>(data2), n_times, n_modes, coef,
        reinterpret_cast<double2*>(LL));
    SCRIB200_CHECK_LAUNCH("ll_comparison");
    return SCRIB200_OK;
}

extern "C" size_t scrib200_dominant_eigenvector_workspace_bytes(int64_t n_times) {
    const size_t nblk = (size_t)(n_times + SIGN_BLK - 1) / SIGN_BLK + 1;
    return (size_t)n_times * 2 + 4 * nblk + 16;
}

extern "C" int scrib200_dominant_eigenvector(const double* LL, int64_t n_times, const double* rough_direction,
                                             int64_t rough_index, double* dpa, void* workspace, size_t workspace_bytes,
                                             void* stream) {
    SCRIB200_REQUIRE(LL && rough_direction && dpa && workspace, "dominant_eigenvector: null pointer");
    SCRIB200_REQUIRE(workspace_bytes >= scrib200_dominant_eigenvector_workspace_bytes(n_times),
                     "dominant_eigenvector: workspace too small");
    if (n_times <= 0) return SCRIB200_OK;
    if (rough_index < 0) rough_index += n_times;
    SCRIB200_REQUIRE(rough_index >= 0 && rough_index < n_times, "dominant_eigenvector: rough_index out of range");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* maps = reinterpret_cast<unsigned char*>(workspace);
    signed char* signs = reinterpret_cast<signed char*>(maps + n_times);
    const unsigned nb = (unsigned)((n_times + 127) / 128);
    sym3_dominant_eig_kernel<<<nb, 128, 0, st>>>(LL, n_times, dpa);
    SCRIB200_CHECK_LAUNCH("dominant_eigenvector(eig)");
    eig_sign_map_kernel<<<nb, 128, 0, st>>>(dpa, n_times, rough_index, maps);
    SCRIB200_CHECK_LAUNCH("dominant_eigenvector(map)");
    const int64_t nblk_dir = (n_times + SIGN_BLK - 1) / SIGN_BLK + 1;   // blocks per direction (upper bound)
    unsigned char* blkmap = reinterpret_cast<unsigned char*>(signs + n_times);
    unsigned char* blkstate = blkmap + 2 * nblk_dir;
    const unsigned nbb = (unsigned)((2 * nblk_dir + 127) / 128);
    eig_sign_blocks_kernel<<<nbb, 128, 0, st>>>(maps, n_times, rough_index, nblk_dir, blkmap);
    SCRIB200_CHECK_LAUNCH("dominant_eigenvector(scan: blocks)");
    eig_sign_walk_kernel<<<1, 256, 2 * 16384, st>>>(dpa, rough_index, rough_direction, nblk_dir, blkmap, blkstate, signs);
    SCRIB200_CHECK_LAUNCH("dominant_eigenvector(scan: walk)");
    eig_sign_fill_kernel<<<nbb, 128, 0, st>>>(maps, n_times, rough_index, nblk_dir, blkstate, signs);
    SCRIB200_CHECK_LAUNCH("dominant_eigenvector(scan: fill)");
    eig_sign_finish_kernel<<<nb, 128, 0, st>>>(dpa, n_times, signs);
    SCRIB200_CHECK_LAUNCH("dominant_eigenvector(finish)");
    return SCRIB200_OK;
}

extern "C" int scrib200_solve3(const double* A, const double* b, int64_t n, double scale, double* x, void* stream) {
    SCRIB200_REQUIRE(A && b && x, "solve3: null pointer");
    if (n <= 0) return SCRIB200_OK;
    solve3_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(A, b, n, scale, x);
    SCRIB200_CHECK_LAUNCH("solve3");
    return SCRIB200_OK;
}

extern "C" int scrib200_sparse_expectation(const double* a, const double* b, int64_t n_times, int n_modes,
                                           const int* rows, const int* cols, const double* vals, const int* seg_host,
                                           const int* seg_dev, int K, double* out, void* stream) {
    SCRIB200_REQUIRE(a && b && rows && cols && vals && seg_dev && out, "sparse_expectation: null pointer");
    SCRIB200_REQUIRE(K >= 1, "sparse_expectation: K=%d", K);
    (void)seg_host;
    if (n_times <= 0) return SCRIB200_OK;
    sparse_expectation_kernel<<<warp_blocks((n_times + SPARSE_ST - 1) / SPARSE_ST), 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2*>(a), reinterpret_cast<const double2*>(b), n_times, n_modes, rows, cols,
        reinterpret_cast<const double2*>(vals), seg_dev, K, reinterpret_cast<double2*>(out));
    SCRIB200_CHECK_LAUNCH("sparse_expectation");
    return SCRIB200_OK;
}

extern "C" int scrib200_sparse_expectation_ell(const double* a, const double* b, int64_t n_times, int n_modes, const int* ell_rows,
                                               const double* ell_vals, const int* widths_host, int K, int real_values, double* out,
                                               void* stream) {
    SCRIB200_REQUIRE(a && b && ell_rows && ell_vals && widths_host && out, "sparse_expectation_ell: null pointer");
    SCRIB200_REQUIRE(K >= 1 && K <= 32, "sparse_expectation_ell: K=%d must be 1..32", K);
    if (n_times <= 0) return SCRIB200_OK;
    EllLayout lay;
    int off = 0;
    for (int k = 0; k < 32; ++k) {
        lay.width[k] = k < K ? widths_host[k] : 0;
        lay.offset[k] = off;
        SCRIB200_REQUIRE(lay.width[k] >= 0 && lay.width[k] <= 64, "sparse_expectation_ell: width %d", lay.width[k]);
        off += lay.width[k] * n_modes;
    }
    const unsigned blocks = warp_blocks(n_times);
    for (int k0 = 0; k0 < K; k0 += 4) {
        if (real_values)
            sparse_expectation_ell_kernel<4, true><<<blocks, 128, 0, (cudaStream_t)stream>>>(
                reinterpret_cast<const double2*>(a), reinterpret_cast<const double2*>(b), n_times, n_modes, ell_rows, ell_vals, lay, k0, K,
                reinterpret_cast<double2*>(out));
        else
            sparse_expectation_ell_kernel<4, false><<<blocks, 128, 0, (cudaStream_t)stream>>>(
                reinterpret_cast<const double2*>(a), reinterpret_cast<const double2*>(b), n_times, n_modes, ell_rows, ell_vals, lay, k0, K,
                reinterpret_cast<double2*>(out));
        SCRIB200_CHECK_LAUNCH("sparse_expectation_ell");
    }
    return SCRIB200_OK;
}

extern "C" int scrib200_sparse_expectation_time(const double* a, const double* b, int64_t n_times, int n_modes, const void* entries,
                                                int width, int K, int complex_values, const int* blocks_host, int n_blocks,
                                                double* out, void* stream) {
    SCRIB200_REQUIRE(a && b && entries && blocks_host && out, "sparse_expectation_time: null pointer");
    SCRIB200_REQUIRE(aligned16(a) && aligned16(b) && aligned16(entries), "sparse_expectation_time: pointers must be 16-byte aligned");
    SCRIB200_REQUIRE(K >= 1 && K <= 32, "sparse_expectation_time: K=%d must be 1..32", K);
    SCRIB200_REQUIRE(width >= 1 && width <= XT_MAXW, "sparse_expectation_time: width %d (1..%d)", width, XT_MAXW);
    SCRIB200_REQUIRE(n_blocks >= 1 && n_blocks <= XT_MAXBLK, "sparse_expectation_time: %d column blocks (1..%d)", n_blocks, XT_MAXBLK);
    if (n_times <= 0) return SCRIB200_OK;
    XtPlan plan;
    plan.n_blocks = n_blocks;
    plan.max_rows = plan.max_cols = 0;
    for (int i = 0; i < XT_MAXBLK; ++i) {
        const bool live = i < n_blocks;
        plan.c0[i] = live ? blocks_host[4 * i] : 0;
        plan.c1[i] = live ? blocks_host[4 * i + 1] : 0;
        plan.rlo[i] = live ? blocks_host[4 * i + 2] : 0;
        plan.rhi[i] = live ? blocks_host[4 * i + 3] : 0;
        if (live) {
            SCRIB200_REQUIRE(0 <= plan.c0[i] && plan.c0[i] < plan.c1[i] && plan.c1[i] <= n_modes && 0 <= plan.rlo[i] && plan.rlo[i] < plan.rhi[i] &&
                                 plan.rhi[i] <= n_modes,
                             "sparse_expectation_time: bad column block %d", i);
            plan.max_rows = std::max(plan.max_rows, plan.rhi[i] - plan.rlo[i]);
            plan.max_cols = std::max(plan.max_cols, plan.c1[i] - plan.c0[i]);
        }
    }
    const bool same = (a == b);
    if (same)      // b[c] is read from the rows staged for a: every block's columns must lie inside its row window
        for (int i = 0; i < n_blocks; ++i)
            SCRIB200_REQUIRE(plan.rlo[i] <= plan.c0[i] && plan.c1[i] <= plan.rhi[i], "sparse_expectation_time: block %d: columns outside the row window", i);
    plan.slots = K * width;
    for (int k = 0; k < 4; ++k) plan.width[k] = width;
    const size_t tile_rows = (size_t)plan.max_rows + (same ? 0 : plan.max_cols);
    const size_t table_words = (size_t)plan.max_cols * plan.slots * (complex_values ? 2 : 1);
    const size_t smem = std::max(tile_rows * XT_PITCH + table_words, (size_t)XT_WARPS * 4 * 32) * sizeof(double2);
    SCRIB200_REQUIRE(smem <= 110 * 1024, "sparse_expectation_time: blocks need %zu bytes of shared memory", smem);
    const unsigned grid = (unsigned)((n_times + XT_T - 1) / XT_T);
    for (int k0 = 0; k0 < K; k0 += 4) {
        plan.slot0 = k0 * width;
#define XT_LAUNCH(W_, CPLX_, SAME_)                                                                                          \
    {                                                                                                                        \
        cudaFuncSetAttribute(sparse_expectation_time_kernel<W_, CPLX_, SAME_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        sparse_expectation_time_kernel<W_, CPLX_, SAME_><<<grid, 32 * XT_WARPS, smem, (cudaStream_t)stream>>>(               \
            reinterpret_cast<const double2*>(a), reinterpret_cast<const double2*>(b), n_times, n_modes,                      \
            reinterpret_cast<const double2*>(entries), plan, k0, K, reinterpret_cast<double2*>(out));                        \
    }
#define XT_PICK(W_)                                                                                                          \
    if (complex_values) {                                                                                                    \
        if (same) XT_LAUNCH(W_, true, true) else XT_LAUNCH(W_, true, false)                                                  \
    } else {                                                                                                                 \
        if (same) XT_LAUNCH(W_, false, true) else XT_LAUNCH(W_, false, false)                                                \
    }
        switch (width) {
            case 1: XT_PICK(1) break;
            case 2: XT_PICK(2) break;
            case 3: XT_PICK(3) break;
            default: XT_PICK(4) break;
        }
#undef XT_PICK
#undef XT_LAUNCH
        SCRIB200_CHECK_LAUNCH("sparse_expectation_time");
    }
    return SCRIB200_OK;
}
