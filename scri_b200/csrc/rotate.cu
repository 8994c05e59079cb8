// K4: per-time-step Wigner-D rotation of the modes, in place.
//
// Replaces scri/rotations.py:346-392 (numba loops) + sf._Wigner_D_matrices (per time step).
// a'_{l m}(t) = sum_{m'} a_{l m'}(t) D^l_{m',m}(R(t)).
//
// Three kernels, newest last (see each one's header):
//   * rotate_modes_kernel (any ell_max; used above 16): D is never materialised, D^l_{m'm} = PhA(m'+m) PhB(m-m')
//     P^l_{m'm}(cos beta) with the phases from per-step tables of powers of Ra, Rb (no trigonometry, regular at the poles) and
//     the real polynomial P advanced in l by the three-term recurrence from an exact seed at l0 = max(|m'|, |m|) (tables:
//     scri_b200/_sf.py:wigner_tables); a CTA owns TB time steps staged through shared memory, one thread per (step, m);
//   * rotate_modes_time_kernel (ell_max <= 16, SCRIB200_ROTATE_RECURRENCE): the same recurrences with lanes along time, four
//     matrix elements per recurrence step by symmetry, a rung of ~22 instructions;
//   * rotate_modes_dmma_kernel (ell_max <= 16, the default): D^l(R) factored into the constant matrices d^l(pi/2) and diagonal
//     phases - two FP64 tensor-core products per (16 time steps, l).
#include <stdlib.h>

#include <mutex>
#include <vector>

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

// ---- time-lane kernel (ell_max <= 16): lanes run along TIME, a warp owns one |m| (and the opposite end hi - |m| of the
// range, for balance).  Everything that steers the m' and l loops - the start l0 = max(m', |m|), the recurrence
// coefficients, the seeds - is then uniform across the warp (read from constant memory, no predication, every lane
// busy), and each thread serves FOUR matrix elements per recurrence step through the symmetries
//     P(-m', -m) = (-1)^{m'+m} P(m', m),      P(m', -m) = (-1)^{m'+m} P(-m', m):
// the chains P = P(m', m) and N = P(-m', m) give   out[+m] += P b[+m'] + N b[-m'],   out[-m] += (-1)^{m'+m} (N b[+m'] + P b[-m']).
// The tile of 32 time steps is staged through shared memory transposed ([mode][time], pitch 33) and premultiplied by
// e^{i m'(A-B)}; the accumulators of one block of l live in registers (ell_max > 8 takes three launches over blocks of l,
// the recurrence restarted), the result is multiplied by e^{i m(A+B)} and stored.  A time step whose Rb is exactly zero
// is finished by an exact diagonal pass so that the identity rotation stays bit-exact.
//
// What one rung of the ladder (one l, four matrix elements) costs decides the kernel - the tensor cores have no part in a
// product whose matrix changes with every time step - so the rung is written down to the instruction:
//   * the l -> l+1 coefficients (a, b, c) of (m', m) come ready-made from a compact constant table (1496 triples, all
//     l <= 16): x = a cos(beta) -+ b, P' = x P - c P'' is 6 FP64 instructions for the two chains;
//   * the sign (-1)^{m'} is a compile-time property of the ladder (the m' loop is unrolled by two), (-1)^m is applied once at
//     the end: the four accumulations are 8 FMAs with no multiplication by a sign;
//   * the two shared-memory operands are read at [register + immediate] (inline PTX: left to itself the compiler
//     recomputed both addresses on every rung to save two registers);
//   * in the EXACT instantiations (the block of l is covered completely: ell_min <= LA, ell_max >= LB) no rung tests anything.
__constant__ double c_rot_seed[33 * 33];      // [m' + 16][m + 16]
__constant__ double c_rot_rec[1496 * 3];      // (a, b, c) of the step l -> l+1 for 0 <= m, m' <= 16, l = max(m, m') .. 15
__constant__ int c_rot_off[17 * 17];          // [m][m']: index of the (m, m', l = max(m, m')) triple

constexpr int ROT3_TB = 32;
constexpr int ROT3_PITCH = ROT3_TB + 1;

// seed table at a fixed pitch: seed[(m' + 16) * 33 + m + 16] - the kernel's table indices do not depend on ell_max
constexpr int ROT3_C = 16, ROT3_NM = 33;

__device__ __forceinline__ double2 rot_lds_reg(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

template <int OFF>
__device__ __forceinline__ double2 rot_lds_imm(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF));
    return v;
}

template <int LA, int LB, bool EXACT>
__device__ __forceinline__ void rot4_pass(unsigned sb_addr, const double2* __restrict__ s_b, const double* __restrict__ s_ra,
                                          const double* __restrict__ s_rb, const double2* __restrict__ s_pw, double cosb, bool diagonal,
                                          int tt, int m, int L, int ell_min, double2* __restrict__ orow) {
    constexpr int NA = LB - LA + 1;
    const int lo = EXACT ? LA : (LA > ell_min ? LA : ell_min);          // the tile holds l = lo .. min(LB, L)
    const int lmin2 = lo * lo, omin2 = ell_min * ell_min;
    double pr[NA], pi[NA], nr[NA], ni[NA];                  // out[+m] and (-1)^m out[-m] for l = LA .. LB
#pragma unroll
    for (int i = 0; i < NA; ++i) pr[i] = pi[i] = nr[i] = ni[i] = 0.0;
    const int lend = EXACT ? LB : (LB < L ? LB : L);
    // seeds P^{l0}(+-m', m) ra^.. rb^.. of the NEXT m' are formed while the ladder of the current one runs (their shared-
    // memory and constant loads would otherwise sit in front of every ladder with only a few warps per scheduler to hide them)
    auto seeds = [&](int mp, double& P, double& N) {
        const int ka = mp + m, kb = mp > m ? mp - m : m - mp;            // exponents of ra, rb for (m', m); swapped for (-m', m)
        const double ra_a = s_ra[ka * ROT3_TB + tt], rb_b = s_rb[kb * ROT3_TB + tt];
        const double ra_b = s_ra[kb * ROT3_TB + tt], rb_a = s_rb[ka * ROT3_TB + tt];
        P = c_rot_seed[(mp + ROT3_C) * ROT3_NM + m + ROT3_C] * (ra_a * rb_b);
        N = (mp > 0) ? c_rot_seed[(ROT3_C - mp) * ROT3_NM + m + ROT3_C] * (ra_b * rb_a) : 0.0;
    };
    double Pn, Nn;
    seeds(0, Pn, Nn);
    // one ladder: m' = mp_, ODD says whether m' is odd (the sign of the out[-m] accumulation)
#define ROT4_OFF(l_) (((l_) * ((l_) + 1) - LB * (LB + 1)) * ROT3_PITCH * 16)
    // Software pipeline: a rung works on operands that are in registers already (set l & 1) and first issues the loads of
    // the next rung into the other set - every rung is a switch label, i.e. a basic block of its own, and left to itself
    // each block would start by waiting for its own shared-memory and constant loads (measured: 6 warps stalled on the
    // short scoreboard per instruction issued).  The entry rung's operands are loaded before the switch, into both sets.
#define ROT4_RUNG(l_, ODD)                                                                                     \
    case l_:                                                                                                   \
        if (l_ <= LB && (EXACT || l_ <= lend)) {                                                               \
            constexpr int cur = (l_) & 1, nxt = cur ^ 1;                                                       \
            if (l_ + 1 <= LB && (EXACT || l_ + 1 <= lend)) {                                                   \
                if (l_ + 1 >= LA && (EXACT || l_ + 1 >= ell_min)) {                                            \
                    bpv[nxt] = rot_lds_imm<ROT4_OFF(l_ + 1)>(bp_addr);                                         \
                    bnv[nxt] = rot_lds_imm<ROT4_OFF(l_ + 1)>(bn_addr);                                         \
                }                                                                                              \
                if (l_ + 1 < LB && (EXACT || l_ + 1 < lend)) {                                                 \
                    cav[nxt] = recp[(l_ + 1) * 3];                                                             \
                    cbv[nxt] = recp[(l_ + 1) * 3 + 1];                                                         \
                    ccv[nxt] = recp[(l_ + 1) * 3 + 2];                                                         \
                }                                                                                              \
            }                                                                                                  \
            if (l_ >= LA && (EXACT || l_ >= ell_min)) {                                                        \
                const double2 bp = bpv[cur], bn = bnv[cur];                                                    \
                constexpr int ia = (l_ >= LA) ? l_ - LA : 0;                                                   \
                pr[ia] = fma(P, bp.x, fma(N, bn.x, pr[ia]));                                                   \
                pi[ia] = fma(P, bp.y, fma(N, bn.y, pi[ia]));                                                   \
                if (ODD) {                                                                                     \
                    nr[ia] = fma(-N, bp.x, fma(-P, bn.x, nr[ia]));                                             \
                    ni[ia] = fma(-N, bp.y, fma(-P, bn.y, ni[ia]));                                             \
                } else {                                                                                       \
                    nr[ia] = fma(N, bp.x, fma(P, bn.x, nr[ia]));                                               \
                    ni[ia] = fma(N, bp.y, fma(P, bn.y, ni[ia]));                                               \
                }                                                                                              \
            }                                                                                                  \
            if (l_ < LB && (EXACT || l_ < lend)) {                                                             \
                const double ca = cav[cur], cb = cbv[cur], cc = ccv[cur];                                      \
                const double Pq = fma(fma(ca, cosb, -cb), P, -(cc * P1));                                      \
                const double Nq = fma(fma(ca, cosb, cb), N, -(cc * N1));                                       \
                P1 = P;                                                                                        \
                N1 = N;                                                                                        \
                P = Pq;                                                                                        \
                N = Nq;                                                                                        \
            }                                                                                                  \
        }
#define ROT4_LADDER(mp_, ODD)                                                                                  \
    {                                                                                                          \
        const int l0 = (mp_) > m ? (mp_) : m;                                                                  \
        double P = Pn, N = Nn;                                                                                 \
        if ((mp_) < lend) seeds((mp_) + 1, Pn, Nn);                                                            \
        double P1 = 0.0, N1 = 0.0;                                                                             \
        /* rows of (LB, +m') and (LB, -m') as shared-memory byte addresses held in registers (the rungs read at */ \
        /* negative immediates from there: a base at l = 0 would lie below the tile)                           */ \
        unsigned bp_addr = sb_addr + (unsigned)(((LB * (LB + 1) - lmin2 + (mp_)) * ROT3_PITCH + tt) * 16);     \
        unsigned bn_addr = sb_addr + (unsigned)(((LB * (LB + 1) - lmin2 - (mp_)) * ROT3_PITCH + tt) * 16);     \
        asm volatile("" : "+r"(bp_addr), "+r"(bn_addr));                                                       \
        int rec_i = 3 * (c_rot_off[m * 17 + (mp_)] - l0);                          /* + 3 l */                 \
        asm volatile("" : "+r"(rec_i));                                                                        \
        const double* recp = c_rot_rec + rec_i;                                                                \
        double2 bpv[2], bnv[2];                                                                                \
        double cav[2], cbv[2], ccv[2];                                                                         \
        {   /* operands of the entry rung */                                                                   \
            const int o0 = (l0 * (l0 + 1) - LB * (LB + 1)) * ROT3_PITCH * 16;                                  \
            bpv[0] = bnv[0] = make_double2(0.0, 0.0);                                                          \
            if (l0 >= lo) {                                                                                    \
                bpv[0] = rot_lds_reg(bp_addr + (unsigned)o0);                                                  \
                bnv[0] = rot_lds_reg(bn_addr + (unsigned)o0);                                                  \
            }                                                                                                  \
            cav[0] = cbv[0] = ccv[0] = 0.0;                                                                    \
            if (l0 < lend) {                                                                                   \
                cav[0] = recp[l0 * 3];                                                                         \
                cbv[0] = recp[l0 * 3 + 1];                                                                     \
                ccv[0] = recp[l0 * 3 + 2];                                                                     \
            }                                                                                                  \
            bpv[1] = bpv[0];                                                                                   \
            bnv[1] = bnv[0];                                                                                   \
            cav[1] = cav[0];                                                                                   \
            cbv[1] = cbv[0];                                                                                   \
            ccv[1] = ccv[0];                                                                                   \
        }                                                                                                      \
        /* the unrolled ladder over l is entered at l0 (uniform across the warp): no test per skipped rung */  \
        switch (l0) {                                                                                          \
            ROT4_RUNG(0, ODD) ROT4_RUNG(1, ODD) ROT4_RUNG(2, ODD) ROT4_RUNG(3, ODD) ROT4_RUNG(4, ODD) ROT4_RUNG(5, ODD)       \
            ROT4_RUNG(6, ODD) ROT4_RUNG(7, ODD) ROT4_RUNG(8, ODD) ROT4_RUNG(9, ODD) ROT4_RUNG(10, ODD) ROT4_RUNG(11, ODD)     \
            ROT4_RUNG(12, ODD) ROT4_RUNG(13, ODD) ROT4_RUNG(14, ODD) ROT4_RUNG(15, ODD) ROT4_RUNG(16, ODD)                    \
            default: break;                                                                                    \
        }                                                                                                      \
    }
    for (int mp = 0; mp <= lend; mp += 2) {                  // l0 = max(m', m) > lend ends the loop: l0 grows with m'
        ROT4_LADDER(mp, false)
        if (mp + 1 <= lend) ROT4_LADDER(mp + 1, true)
    }
#undef ROT4_LADDER
#undef ROT4_RUNG
#undef ROT4_OFF
    const double sgm = (m & 1) ? -1.0 : 1.0;                // the (-1)^m left out of the out[-m] sums
    if (diagonal) {   // Rb == 0 exactly: D^l_{m'm} = delta_{m'm} ra^{2|m|} (phases applied outside)
        const double mag = s_ra[2 * m * ROT3_TB + tt];
#pragma unroll
        for (int l = LA; l <= LB; ++l)
            if (l >= m && l >= ell_min && l <= L) {
                const double2 bp = s_b[(l * (l + 1) - lmin2 + m) * ROT3_PITCH + tt];
                const double2 bn = s_b[(l * (l + 1) - lmin2 - m) * ROT3_PITCH + tt];
                pr[l - LA] = mag * bp.x;
                pi[l - LA] = mag * bp.y;
                nr[l - LA] = sgm * mag * bn.x;
                ni[l - LA] = sgm * mag * bn.y;
            }
    }
    const double2 pw = s_pw[m * ROT3_TB + tt];
    const double2 pwn = cscale(sgm, cconj(pw));
#pragma unroll
    for (int l = LA; l <= LB; ++l)
        if (l >= m && l >= ell_min && l <= L) {
            orow[l * (l + 1) - omin2 + m] = cmul(make_double2(pr[l - LA], pi[l - LA]), pw);
            if (m > 0) orow[l * (l + 1) - omin2 - m] = cmul(make_double2(nr[l - LA], ni[l - LA]), pwn);
        }
}

template <int LA, int LB, bool EXACT>
__global__ void __launch_bounds__(32 * (LB / 2 + 1))
rotate_modes_time_kernel(double2* __restrict__ data, int64_t n_times, int ell_min, int ell_max,
                         const double2* __restrict__ spinors, int64_t spinor_stride) {
    extern __shared__ double2 sm3[];
    const int L = ell_max;
    const int n_modes = L * (L + 2) - ell_min * ell_min + 1;
    const int lo = LA > ell_min ? LA : ell_min, hi = LB < L ? LB : L;   // this launch rotates l = lo .. hi
    const int n_blk = (hi + 1) * (hi + 1) - lo * lo, col0 = lo * lo - ell_min * ell_min;
    double2* s_b = sm3;                                     // [n_blk][33]     a_{l m'} e^{i m'(A-B)}, time along the lanes
    double2* s_pw = s_b + (size_t)n_blk * ROT3_PITCH;       // [hi+1][32]      e^{i k (A+B)}
    double2* s_pu = s_pw + (size_t)(hi + 1) * ROT3_TB;      // [hi+1][32]      e^{i k (A-B)}
    double* s_ra = reinterpret_cast<double*>(s_pu + (size_t)(hi + 1) * ROT3_TB);   // [2 hi + 1][32] ra^k
    double* s_rb = s_ra + (size_t)(2 * hi + 1) * ROT3_TB;   // [2 hi + 1][32]
    double* s_cos = s_rb + (size_t)(2 * hi + 1) * ROT3_TB;  // [32]  cos(beta), or 2 if Rb == 0 exactly
    const int64_t t0 = (int64_t)blockIdx.x * ROT3_TB;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    if (tid < ROT3_TB) {
        const int tt = tid;
        int64_t t = t0 + tt;
        if (t >= n_times) t = n_times - 1;
        const double2 Ra = spinors[t * spinor_stride + 0], Rb = spinors[t * spinor_stride + 1];
        const double ra2 = Ra.x * Ra.x + Ra.y * Ra.y, rb2 = Rb.x * Rb.x + Rb.y * Rb.y;
        const double n2 = ra2 + rb2;
        const double ra = sqrt(ra2 / n2), rb = sqrt(rb2 / n2);
        // unit phases; the phase of an exact zero is 1 (its magnitude powers kill every term it would enter)
        const double2 ea = ra2 > 0.0 ? cscale(1.0 / sqrt(ra2), Ra) : make_double2(1.0, 0.0);
        const double2 eb = rb2 > 0.0 ? cscale(1.0 / sqrt(rb2), Rb) : make_double2(1.0, 0.0);
        const double2 u = cmul(ea, cconj(eb)), w = cmul(ea, eb);
        double2 pu = make_double2(1.0, 0.0), pw = pu;
        for (int k = 0; k <= hi; ++k) {
            s_pu[k * ROT3_TB + tt] = pu;
            s_pw[k * ROT3_TB + tt] = pw;
            pu = cmul(pu, u);
            pw = cmul(pw, w);
        }
        double pa = 1.0, pb = 1.0;
        for (int k = 0; k <= 2 * hi; ++k) {
            s_ra[k * ROT3_TB + tt] = pa;
            s_rb[k * ROT3_TB + tt] = pb;
            pa *= ra;
            pb *= rb;
        }
        s_cos[tt] = (rb2 == 0.0) ? 2.0 : (ra2 - rb2) / n2;
    }
    __syncthreads();
    // stage the l-block of the tile transposed, premultiplied by e^{i m'(A-B)}: a thread keeps its mode (lanes along the
    // modes: coalesced rows of the global read) and walks the 32 time steps - the (l, m') of the mode is decoded once
    const int nt = (n_times - t0 < ROT3_TB) ? (int)(n_times - t0) : ROT3_TB;
    for (int lm = tid; lm < n_blk; lm += nthreads) {
        const int full = lm + lo * lo;
        int l = (int)sqrt((double)full);
        while (l * l > full) --l;
        while ((l + 1) * (l + 1) <= full) ++l;
        const int mp = full - l * (l + 1);
        const double2* ph = s_pu + (mp < 0 ? -mp : mp) * ROT3_TB;
        const double sg = mp < 0 ? -1.0 : 1.0;
        const double2* src = data + t0 * n_modes + col0 + lm;
        double2* dst = s_b + lm * ROT3_PITCH;
#pragma unroll 4
        for (int tt = 0; tt < nt; ++tt) {
            const double2 p = ph[tt];
            dst[tt] = cmul(src[(size_t)tt * n_modes], make_double2(p.x, sg * p.y));
        }
        for (int tt = nt; tt < ROT3_TB; ++tt) dst[tt] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    const int tt = tid & 31, w = tid >> 5;
    if (t0 + tt >= n_times) return;
    const double cosb = s_cos[tt];
    const bool diagonal = cosb > 1.5;
    double2* orow = data + (t0 + tt) * n_modes;
    const unsigned sb_addr = (unsigned)__cvta_generic_to_shared(s_b);
    // warp w owns m = w and m = hi - w (one of them when they coincide)
    for (int which = 0; which < 2; ++which) {
        const int m = which == 0 ? w : hi - w;
        if (m > hi || m < 0 || (which == 1 && m <= w)) continue;
        rot4_pass<LA, LB, EXACT>(sb_addr, s_b, s_ra, s_rb, s_pw, cosb, diagonal, tt, m, L, ell_min, orow);
    }
}

// ---- tensor-core kernel (ell_max <= 16).  The matrix of a rotation changes with every time step, so applying it is not a
// GEMM - but it factors into constant matrices and diagonal phases.  With Ra = ra e^{iA}, Rb = rb e^{iB},
// cos(beta) = ra^2 - rb^2, sin(beta) = 2 ra rb and Delta^l = d^l(pi/2) (real, orthogonal, the same for every time step)
//     D^l_{m'm}(R) = e^{i m'(A-B)} i^{-m'} [ sum_mu Delta_{m' mu} e^{-i mu beta} Delta_{m mu} ] i^{m} e^{i m(A+B)},
// i.e. for the 16 time steps of a tile and one l:   x = a (-i e^{i(A-B)})^{m'}   (diagonal),   y = Delta^T x   (DMMA),
// z = y e^{-i mu beta}   (diagonal),   o = Delta z   (DMMA),   a' = o (i e^{i(A+B)})^m   (diagonal).
// 8 (2l+1)^2 flops per mode-vector instead of a recurrence per matrix element, on the FP64 tensor cores, and nothing
// depends on beta being away from the poles.  One CTA owns 16 time steps; its warps take the values of l from a queue
// (largest first), each staging its [16 x (2l+1)] slab in its own shared-memory tile (time-major, pitch 88 doubles: the
// B fragments of m8n8k4 - lane (k, n) reads mode 4kk + k of step n/2, re or im - touch 32 distinct 8-byte words).  The
// complex pair of one (mode, step) lands in one thread's accumulator pair, so the diagonal steps are register work.
// A time step whose Rb is exactly zero is finished at load time by its (exact) diagonal so that the identity rotation stays
// bit-exact: every factor is then a power of +-i or 1.
constexpr int RG_T = 16;          // time steps per work item (one warp, one l): 4 n-tiles of 8 real columns
constexpr int RG_NT = 4;
constexpr int RG_WARPS = 8;
struct RotGemmOffsets {
    int off[17];                  // start (in doubles) of the A fragments of l: Delta^T tiles [Kt][Mt][32], then Delta tiles
};

__device__ __forceinline__ void rot_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__host__ __device__ inline int rot_phase_pitch(int L) { return ((L + 1 + 5) & ~7) + 2; }   // >= L + 1, = 2 (mod 8)

// One product of the pair: acc[mm][nn] = sum_kk A(mm, kk) B(kk, nn) for the MA m-tiles this l needs (compile time: no
// predicated-off DMMAs, no per-step tests).  A fragments [kk][mm][lane] from the packed table, fetched one k-step ahead
// (they come from L2 more often than not); B fragments from the tile at [register + immediate].
template <int MT, int MA, int PITCH>
__device__ __forceinline__ void rot_gemm_pass(double (&acc)[MT][RG_NT][2], const double* __restrict__ af, const double* __restrict__ bp, int Kt) {
#pragma unroll
    for (int mm = 0; mm < MA; ++mm)
#pragma unroll
        for (int nn = 0; nn < RG_NT; ++nn) acc[mm][nn][0] = acc[mm][nn][1] = 0.0;
    double a_nxt[MA];
#pragma unroll
    for (int mm = 0; mm < MA; ++mm) a_nxt[mm] = __ldg(af + mm * 32);
    for (int kk = 0; kk < Kt; ++kk) {
        double a_cur[MA], b[RG_NT];
#pragma unroll
        for (int mm = 0; mm < MA; ++mm) a_cur[mm] = a_nxt[mm];
        af += MA * 32;
        if (kk + 1 < Kt) {
#pragma unroll
            for (int mm = 0; mm < MA; ++mm) a_nxt[mm] = __ldg(af + mm * 32);
        }
#pragma unroll
        for (int nn = 0; nn < RG_NT; ++nn) b[nn] = bp[4 * nn * PITCH];
        bp += 8;
#pragma unroll
        for (int mm = 0; mm < MA; ++mm)
#pragma unroll
            for (int nn = 0; nn < RG_NT; ++nn) rot_dmma(acc[mm][nn][0], acc[mm][nn][1], a_cur[mm], b[nn]);
    }
}

// accumulators -> tile, multiplied by the diagonal phase of their row (table[t][|k|], conjugated for k < 0)
template <int MT, int MA, int PITCH>
__device__ __forceinline__ void rot_phase_store(const double (&acc)[MT][RG_NT][2], double* __restrict__ tile, const double2* __restrict__ table,
                                                int PK, int ts0, int l, int n, int g, int q) {
#pragma unroll
    for (int mm = 0; mm < MA; ++mm) {
        const int row = 8 * mm + g, k = row - l, ak = k < 0 ? -k : k;
        const bool live = row < n;
#pragma unroll
        for (int nn = 0; nn < RG_NT; ++nn) {
            const int t = 4 * nn + q;
            double2 v = make_double2(acc[mm][nn][0], acc[mm][nn][1]);
            if (live) {
                double2 p = table[(ts0 + t) * PK + ak];
                if (k < 0) p.y = -p.y;
                v = cmul(v, p);
            }
            *reinterpret_cast<double2*>(tile + t * PITCH + 2 * row) = v;
        }
    }
}

template <int MT, int MA, int PITCH>
__device__ __forceinline__ void rot_two_products(double* __restrict__ tile, const double* __restrict__ af, const double2* __restrict__ s_eb,
                                                 const double2* __restrict__ s_q3, int PK, int ts0, int l, int n, int Kt, int lane) {
    if constexpr (MA <= MT) {
        double acc[MT][RG_NT][2];
        const int g = lane >> 2, q = lane & 3;
        const double* bp = tile + (lane >> 3) * PITCH + 2 * q + ((lane >> 2) & 1);   // B fragment: column n = lane / 4 -> step n / 2, re / im
        rot_gemm_pass<MT, MA, PITCH>(acc, af, bp, Kt);                                // y = Delta^T x
        __syncwarp();                                  // every lane has read its B fragments: the tile can be overwritten
        rot_phase_store<MT, MA, PITCH>(acc, tile, s_eb, PK, ts0, l, n, g, q);         // z = y e^{-i mu beta}
        __syncwarp();
        rot_gemm_pass<MT, MA, PITCH>(acc, af + MA * Kt * 32, bp, Kt);                 // o = Delta z
        __syncwarp();
        rot_phase_store<MT, MA, PITCH>(acc, tile, s_q3, PK, ts0, l, n, g, q);         // a' = o (i e^{i(A+B)})^m
        __syncwarp();
    }
}

// MT: m-tiles the accumulators are sized for (3: ell_max <= 11, 5: ell_max <= 16); S (run time): sub-tiles of 16 time steps
// per CTA - the warps take (sub-tile, l) pairs from a queue, largest l first, so that a short ell range still fills 8 warps.
template <int MT>
__global__ void __launch_bounds__(32 * RG_WARPS, 2)
rotate_modes_dmma_kernel(double2* __restrict__ data, int64_t n_times, int ell_min, int ell_max, const double2* __restrict__ spinors,
                         int64_t spinor_stride, const double* __restrict__ frags, const __grid_constant__ RotGemmOffsets offs, int S) {
    constexpr int PITCH = 16 * MT + 8;                 // doubles per time row of a tile: 88 / 56 (odd multiples of 8)
    extern __shared__ double2 smg[];
    const int L = ell_max, TS = RG_T * S;              // time steps of this CTA
    // phase tables [time step][k], k = 0 .. L at a pitch = 2 (mod 8) complex numbers: a quarter-warp of the accumulator
    // layout (2 adjacent k x 4 time steps) and of the staging loop (8 adjacent k) each read 128 distinct bytes
    const int PK = rot_phase_pitch(L);
    double2* s_q1 = smg;                               // (-i e^{i(A-B)})^k
    double2* s_eb = s_q1 + PK * TS;                    // e^{-i k beta}
    double2* s_q3 = s_eb + PK * TS;                    // (i e^{i(A+B)})^k
    double* s_tiles = reinterpret_cast<double*>(s_q3 + PK * TS);        // [warps][16][PITCH]
    __shared__ int s_next;
    __shared__ int s_diag[RG_T * 8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_modes = L * (L + 2) - ell_min * ell_min + 1;
    const int64_t t0 = (int64_t)blockIdx.x * TS;
    if (tid == 0) s_next = 0;
    // the three phase tables: thread (table, time step)
    for (int id = tid; id < 3 * TS; id += (int)blockDim.x) {
        const int which = id / TS, tt = id - which * TS;
        int64_t t = t0 + tt;
        if (t >= n_times) t = n_times - 1;
        const double2 Ra = spinors[t * spinor_stride + 0], Rb = spinors[t * spinor_stride + 1];
        const double ra2 = Ra.x * Ra.x + Ra.y * Ra.y, rb2 = Rb.x * Rb.x + Rb.y * Rb.y;
        const double n2 = ra2 + rb2;
        // unit phases; the phase of an exact zero is 1 (then beta is 0 or pi and the other angle carries the rotation)
        const double2 ea = ra2 > 0.0 ? cscale(1.0 / sqrt(ra2), Ra) : make_double2(1.0, 0.0);
        const double2 eb = rb2 > 0.0 ? cscale(1.0 / sqrt(rb2), Rb) : make_double2(1.0, 0.0);
        double2 base;
        double2* table;
        if (which == 0) {
            base = cmul(make_double2(0.0, -1.0), cmul(ea, cconj(eb)));
            table = s_q1;
            s_diag[tt] = (rb2 == 0.0);
        } else if (which == 1) {
            const double ra = sqrt(ra2 / n2), rb = sqrt(rb2 / n2);
            base = make_double2(ra * ra - rb * rb, -2.0 * ra * rb);
            table = s_eb;
        } else {
            base = cmul(make_double2(0.0, 1.0), cmul(ea, eb));
            table = s_q3;
        }
        double2 p = make_double2(1.0, 0.0);
        for (int k = 0; k <= L; ++k) {
            table[tt * PK + k] = p;
            p = cmul(p, base);
        }
    }
    __syncthreads();
    double* tile = s_tiles + (size_t)warp * RG_T * PITCH;
    const int n_ell = L - ell_min + 1;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(&s_next, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_ell * S) break;
        const int l = L - item / S, sub = item - (item / S) * S;       // largest l first, its sub-tiles next to each other
        const int ts0 = sub * RG_T;                                    // first time step of the sub-tile within the CTA
        if (t0 + ts0 >= n_times) continue;
        const int n = 2 * l + 1, Mt = (n + 7) >> 3, Kt = (n + 3) >> 2;
        const int col_l = l * l - ell_min * ell_min;
        // (a) stage the slab, premultiplied by (-i e^{i(A-B)})^{m'} - eight loads in flight per lane; zero the padding
        // columns.  A lane walks the flattened (time step, mode) index in strides of 32 without dividing.
        const int q32 = 32 / n, r32 = 32 - q32 * n;
        const int rows = (n_times - (t0 + ts0) < RG_T) ? (int)(n_times - (t0 + ts0)) : RG_T;      // time steps that exist
        double2* gbase = data + (t0 + ts0) * n_modes + col_l;
        {
            int t = lane / n, j = lane - t * n;
            while (t < RG_T) {
                double2 v[8];
                int tu[8], ju[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    tu[u] = t;
                    ju[u] = j;
                    v[u] = make_double2(0.0, 0.0);
                    if (t < rows) v[u] = gbase[(size_t)t * n_modes + j];
                    t += q32;
                    j += r32;
                    if (j >= n) {
                        j -= n;
                        ++t;
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (tu[u] >= RG_T) continue;
                    double2 x = v[u];
                    if (tu[u] < rows) {
                        const int k = ju[u] - l, ak = k < 0 ? -k : k;
                        double2 p = s_q1[(ts0 + tu[u]) * PK + ak];
                        if (k < 0) p.y = -p.y;
                        x = cmul(x, p);
                        if (s_diag[ts0 + tu[u]]) {
                            double2 p3 = s_q3[(ts0 + tu[u]) * PK + ak];
                            if (k < 0) p3.y = -p3.y;
                            gbase[(size_t)tu[u] * n_modes + ju[u]] = cmul(x, p3);
                        }
                    }
                    *reinterpret_cast<double2*>(tile + tu[u] * PITCH + 2 * ju[u]) = x;
                }
            }
        }
        const int npad = 8 * Mt - n;
        for (int idx = lane; idx < RG_T * npad; idx += 32) {
            const int t = idx / npad, j = n + idx - t * npad;
            *reinterpret_cast<double2*>(tile + t * PITCH + 2 * j) = make_double2(0.0, 0.0);
        }
        __syncwarp();
        const double* af = frags + offs.off[l] + lane;
        switch (Mt) {
            case 1: rot_two_products<MT, 1, PITCH>(tile, af, s_eb, s_q3, PK, ts0, l, n, Kt, lane); break;
            case 2: rot_two_products<MT, 2, PITCH>(tile, af, s_eb, s_q3, PK, ts0, l, n, Kt, lane); break;
            case 3: rot_two_products<MT, 3, PITCH>(tile, af, s_eb, s_q3, PK, ts0, l, n, Kt, lane); break;
            case 4: rot_two_products<MT, 4, PITCH>(tile, af, s_eb, s_q3, PK, ts0, l, n, Kt, lane); break;
            default: rot_two_products<MT, 5, PITCH>(tile, af, s_eb, s_q3, PK, ts0, l, n, Kt, lane); break;
        }
        // (g) rows back to global memory (steps finished by their diagonal excepted)
        {
            int t = lane / n, j = lane - t * n;
            while (t < rows) {
                if (!s_diag[ts0 + t]) gbase[(size_t)t * n_modes + j] = *reinterpret_cast<const double2*>(tile + t * PITCH + 2 * j);
                t += q32;
                j += r32;
                if (j >= n) {
                    j -= n;
                    ++t;
                }
            }
        }
        __syncwarp();
    }
}

static size_t rotate_dmma_smem(int L, int S, int MT, int warps) {
    return 3 * (size_t)rot_phase_pitch(L) * RG_T * S * sizeof(double2) + (size_t)warps * RG_T * (16 * MT + 8) * sizeof(double);
}


static size_t rotate_time_smem(int lo, int hi) {
    const size_t n_blk = (size_t)(hi + 1) * (hi + 1) - (size_t)lo * lo;
    return (n_blk * ROT3_PITCH + 2 * (size_t)(hi + 1) * ROT3_TB) * sizeof(double2) + (2 * (size_t)(2 * hi + 1) * ROT3_TB + ROT3_TB) * sizeof(double);
}

// (a, b, c) of the step l -> l+1, P^{l+1}(m', m) = (a cos(beta) - b) P^l - c P^{l-1}  [+ b for (-m', m)], with
// a = U_l(m') U_l(m), c = V_l(m') V_l(m), b = a m' m / (l (l + 1)) - the factors of scri_b200/_sf.py:wigner_factor_table,
// multiplied out here so that a rung reads three numbers instead of forming them.  Compact: for every 0 <= m, m' <= 16
// the l = max(m, m') .. 15 are contiguous; c_rot_off[m][m'] is where they start.  Built once, uploaded once per device.
static int upload_rotation_recurrence(cudaStream_t st) {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return SCRIB200_OK;
    std::vector<double> rec;
    std::vector<int> off(17 * 17, 0);
    auto U = [](int l, int m) { return sqrt((double)((2 * l + 1) * (l + 1)) / (double)((l + 1) * (l + 1) - m * m)); };
    auto V = [](int l, int m) {
        return l > 0 ? sqrt((double)((l + 1) * (l * l - m * m)) / (double)(l * ((l + 1) * (l + 1) - m * m))) : 0.0;
    };
    for (int m = 0; m <= 16; ++m)
        for (int mp = 0; mp <= 16; ++mp) {
            off[m * 17 + mp] = (int)(rec.size() / 3);
            for (int l = (m > mp ? m : mp); l < 16; ++l) {
                const double a = U(l, mp) * U(l, m);
                const double tb = l > 0 ? (double)(m * mp) * (1.0 / (double)(l * (l + 1))) : 0.0;
                rec.push_back(a);
                rec.push_back(a * tb);
                rec.push_back(V(l, mp) * V(l, m));
            }
        }
    SCRIB200_REQUIRE(rec.size() == 1496 * 3, "rotate_modes: recurrence table has %zu entries", rec.size());
    cudaError_t e = cudaMemcpyToSymbolAsync(c_rot_rec, rec.data(), rec.size() * sizeof(double), 0, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_rot_off, off.data(), off.size() * sizeof(int), 0, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);          // the host vectors go out of scope
    SCRIB200_REQUIRE(e == cudaSuccess, "rotate_modes: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return SCRIB200_OK;
}

}  // namespace scrib200

extern "C" int scrib200_rotate_modes_dmma(double* data, int64_t n_times, int ell_min, int ell_max, const double* spinors,
                                          int64_t spinor_stride, const double* frags, const int* frag_offsets_host, void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(data && spinors && frags && frag_offsets_host, "rotate_modes_dmma: null pointer");
    SCRIB200_REQUIRE(ell_min >= 0 && ell_max >= ell_min && ell_max <= 16, "rotate_modes_dmma: bad ell range [%d, %d] (ell_max <= 16)", ell_min, ell_max);
    SCRIB200_REQUIRE(spinor_stride == 0 || spinor_stride == 2, "rotate_modes_dmma: spinor_stride must be 0 or 2");
    SCRIB200_REQUIRE(aligned16(data) && aligned16(spinors) && aligned16(frags), "rotate_modes_dmma: pointers must be 16-byte aligned");
    if (n_times <= 0) return SCRIB200_OK;
    RotGemmOffsets offs;
    for (int l = 0; l < 17; ++l) offs.off[l] = l <= ell_max ? frag_offsets_host[l] : 0;
    const int n_ell = ell_max - ell_min + 1;
    const int MT = ell_max <= 11 ? 3 : 5;
    // (sub-tiles per CTA, warps): the items of a CTA are few and of very different size (l = 16 costs 20 x l = 2), so one
    // sub-tile per CTA leaves the warps that drew small items idle until the largest is done (15 items on 8 warps: 70 %);
    // two sub-tiles on 7 warps pack to 97 % and still fit two CTAs per SM
    int S = MT == 5 ? 2 : (n_ell >= 6 ? 4 : 8), warps = MT == 5 ? 7 : RG_WARPS;       // measured (1e6 x 285: S, warps = 1, 8: 6.54 ms; 2, 7: 6.09; 3, 6: 6.21; 4, 6: 6.76)
    static const int env_s = [] { const char* e = getenv("SCRIB200_ROTATE_SUBTILES"); return e ? atoi(e) : 0; }();
    static const int env_w = [] { const char* e = getenv("SCRIB200_ROTATE_WARPS"); return e ? atoi(e) : 0; }();
    if (env_s > 0) S = env_s > 8 ? 8 : env_s;
    if (env_w > 0) warps = env_w > (MT == 5 ? 7 : RG_WARPS) ? (MT == 5 ? 7 : RG_WARPS) : env_w;
    while (warps > 4 && rotate_dmma_smem(ell_max, S, MT, warps) > 113 * 1024) --warps;
    while (S > 1 && rotate_dmma_smem(ell_max, S, MT, warps) > 113 * 1024) S /= 2;
    const size_t smem = rotate_dmma_smem(ell_max, S, MT, warps);
    const int64_t blocks = (n_times + RG_T * S - 1) / (RG_T * S);
    SCRIB200_REQUIRE(blocks < (int64_t)2147483647, "rotate_modes_dmma: too many time steps");
    if (MT == 3) {
        cudaFuncSetAttribute(rotate_modes_dmma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        rotate_modes_dmma_kernel<3><<<(unsigned)blocks, 32 * warps, smem, (cudaStream_t)stream>>>(
            reinterpret_cast<double2*>(data), n_times, ell_min, ell_max, reinterpret_cast<const double2*>(spinors), spinor_stride, frags, offs, S);
    } else {
        cudaFuncSetAttribute(rotate_modes_dmma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        rotate_modes_dmma_kernel<5><<<(unsigned)blocks, 32 * warps, smem, (cudaStream_t)stream>>>(
            reinterpret_cast<double2*>(data), n_times, ell_min, ell_max, reinterpret_cast<const double2*>(spinors), spinor_stride, frags, offs, S);
    }
    SCRIB200_CHECK_LAUNCH("rotate_modes_dmma");
    return SCRIB200_OK;
}

extern "C" int scrib200_rotate_modes(double* data, int64_t n_times, int ell_min, int ell_max, const double* spinors,
                                     int64_t spinor_stride, const double* seed, const double* rec, const double* uv,
                                     void* stream) {
    using namespace scrib200;
    SCRIB200_REQUIRE(data && spinors && seed && (rec || uv), "rotate_modes: null pointer");
    SCRIB200_REQUIRE(ell_min >= 0 && ell_max >= ell_min, "rotate_modes: bad ell range [%d, %d]", ell_min, ell_max);
    SCRIB200_REQUIRE(spinor_stride == 0 || spinor_stride == 2, "rotate_modes: spinor_stride must be 0 or 2");
    SCRIB200_REQUIRE(aligned16(data) && aligned16(spinors), "rotate_modes: pointers must be 16-byte aligned");
    if (n_times <= 0) return SCRIB200_OK;
    const int L = ell_max, nm = 2 * L + 1;
    const int n_modes = L * (L + 2) - ell_min * ell_min + 1;
    if (uv && L <= 16 && getenv("SCRIB200_ROTATE_V1") == nullptr) {
        cudaStream_t st = (cudaStream_t)stream;
        // the small tables steer warp-uniform loops: constant memory (broadcast, no load/store-unit traffic).  The seeds are
        // re-laid out at pitch 33 around index 16, so that the kernel's table indices do not depend on ell_max; the
        // recurrence triples are the same for every ell_max <= 16 and go up once per device
        int rc = upload_rotation_recurrence(st);
        if (rc != SCRIB200_OK) return rc;
        void* p_seed = nullptr;
        cudaError_t e = cudaGetSymbolAddress(&p_seed, c_rot_seed);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(reinterpret_cast<double*>(p_seed) + (16 - L) * 33 + (16 - L), 33 * sizeof(double), seed, nm * sizeof(double),
                                  nm * sizeof(double), nm, cudaMemcpyDeviceToDevice, st);
        SCRIB200_REQUIRE(e == cudaSuccess, "rotate_modes: %s", cudaGetErrorString(e));
        const int64_t blocks = (n_times + ROT3_TB - 1) / ROT3_TB;
        // one launch per block of l (0..8, 9..12, 13..16): each stages only its own modes (39 / 47 / 63 KB of shared memory
        // instead of 150 KB: 5, 3 and 2 CTAs per SM), reads and writes its own columns of `data`, restarts the recurrences at
        // l0.  A block that is covered completely takes the instantiation without per-rung tests (LA = its first l).
#define ROT4_LAUNCH(LA_, LB_, EX_, lo_, hi_)                                                                                      \
    {                                                                                                                             \
        const size_t smem = rotate_time_smem(lo_, hi_);                                                                           \
        cudaFuncSetAttribute(rotate_modes_time_kernel<LA_, LB_, EX_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        rotate_modes_time_kernel<LA_, LB_, EX_><<<(unsigned)blocks, 32 * ((hi_) / 2 + 1), smem, st>>>(                            \
            reinterpret_cast<double2*>(data), n_times, ell_min, ell_max, reinterpret_cast<const double2*>(spinors), spinor_stride); \
        SCRIB200_CHECK_LAUNCH("rotate_modes");                                                                                    \
    }
#define ROT4_BLOCK(LA_, LB_)                                                                                                      \
    if (ell_min <= LB_ && L >= LA_) {                                                                                             \
        const int lo = ell_min > LA_ ? ell_min : LA_, hi = L < LB_ ? L : LB_;                                                     \
        ROT4_LAUNCH(LA_, LB_, false, lo, hi)                                                                                      \
    }
        if (ell_min == 2 && L >= 8) ROT4_LAUNCH(2, 8, true, 2, 8) else ROT4_BLOCK(0, 8)
        if (ell_min <= 9 && L >= 12) ROT4_LAUNCH(9, 12, true, 9, 12) else ROT4_BLOCK(9, 12)
        if (ell_min <= 13 && L >= 16) ROT4_LAUNCH(13, 16, true, 13, 16) else ROT4_BLOCK(13, 16)
#undef ROT4_BLOCK
#undef ROT4_LAUNCH
        return SCRIB200_OK;
    }
    SCRIB200_REQUIRE(rec, "rotate_modes: the recurrence table `rec` is needed for ell_max > 16");
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
