"""Host-side special-function tables for the B200 path (plan construction only).

These are the product's own O(G*L^2) table builders; the per-time-step work happens on the GPU
(scri_b200/csrc).  They replace, for this path, the pieces of `spherical_functions` / `spinsfast`
that the reference calls at scri/waveform_grid.py:470-471 (sf.SWSH_grid), rotations.py:299,327
(Wigner-D storage layout) and waveform_grid.py:305 (spinsfast.map2salm quadrature).

Algorithm (not the reference's): Wigner D^l_{m',m}(R) = PhA(m'+m) PhB(m-m') P^l_{m'm}(|Ra|^2,|Rb|^2)
with PhA(k)=Ra^k (k>=0) or conj(Ra)^|k|, PhB likewise for Rb, and the real polynomial P advanced
in l by the three-term recurrence of Wigner's d from an exact single-term seed at
l0=max(|m'|,|m|).  No trigonometry, no division by |Ra| or |Rb| - regular at both poles - and
stable (upward recurrence follows the dominant solution).  The same coefficient tables are
uploaded to the GPU for the per-time-step rotation kernel.
"""
import math
from functools import lru_cache

import numpy as np


# ----------------------------------------------------------------------------- (l,m) layout
def LM_index(ell, m, ell_min):
    """Flat index of (ell, m): ell(ell+1) - ell_min^2 + m  (scri/waveform_modes.py:455)."""
    return ell * (ell + 1) - ell_min**2 + m


def LM_total_size(ell_min, ell_max):
    return ell_max * (ell_max + 2) - ell_min**2 + 1


def LM_range(ell_min, ell_max):
    return np.array([[ell, m] for ell in range(ell_min, ell_max + 1) for m in range(-ell, ell + 1)], dtype=np.int64)


def D_offset(ell, ell_min):
    """Start of the ell block in the flat Wigner-D array (rotations.py:359,384 use sf._linear_matrix_offset)."""
    return ((4 * ell * ell - 1) * ell - (4 * ell_min * ell_min - 1) * ell_min) // 3


def D_total_size(ell_min, ell_max):
    return D_offset(ell_max + 1, ell_min)


# ----------------------------------------------------------------------------- recurrence tables
def _seed(mp, m):
    """P^{l0}_{mp,m}: exact single-term value at l0 = max(|mp|,|m|)."""
    ell = max(abs(mp), abs(m))
    num = math.factorial(ell + m) * math.factorial(ell - m)
    den = math.factorial(ell + mp) * math.factorial(ell - mp)
    tot = 0
    for rho in range(max(0, mp - m), min(ell + mp, ell - m) + 1):
        tot += (-1) ** rho * math.comb(ell + mp, rho) * math.comb(ell - mp, ell - rho - m)
    g = math.gcd(num, den)
    return tot * math.sqrt(num // g) / math.sqrt(den // g)


def _rec_coeffs(ell, mp, m):
    """(a,b,c) with P^{l+1} = (a*cos(beta) - b) P^l - c P^{l-1}."""
    if ell == 0:
        return 1.0, 0.0, 0.0
    den = math.sqrt(((ell + 1) ** 2 - mp * mp) * ((ell + 1) ** 2 - m * m))
    a = (2 * ell + 1) * (ell + 1) / den
    b = (2 * ell + 1) * mp * m / (ell * den)
    c = (ell + 1) * math.sqrt((ell * ell - mp * mp) * (ell * ell - m * m)) / (ell * den)
    return a, b, c


@lru_cache(maxsize=None)
def wigner_tables(ell_max):
    """Tables shared by host and device.

    seed[mp+L, m+L]            : P at l0
    rec[l, mp+L, m+L, 0:3]     : (a,b,c) to step l -> l+1   (l in [0, L-1]); zero where l < l0
    """
    L = ell_max
    n = 2 * L + 1
    seed = np.zeros((n, n))
    rec = np.zeros((max(L, 1), n, n, 3))
    for mp in range(-L, L + 1):
        for m in range(-L, L + 1):
            seed[mp + L, m + L] = _seed(mp, m)
            for ell in range(max(abs(mp), abs(m)), L):
                rec[ell, mp + L, m + L] = _rec_coeffs(ell, mp, m)
    return seed, rec


@lru_cache(maxsize=None)
def wigner_factor_table(ell_max):
    """uv[l, m+L] = (U, V) with the l -> l+1 recurrence coefficients a = U[l][m']U[l][m], c = V[l][m']V[l][m] and
    b = a m'm/(l(l+1))  (the same coefficients as `_rec_coeffs`, factored so that the table is O(L^2))."""
    L = ell_max
    out = np.zeros((max(L, 1), 2 * L + 1, 2))
    for ell in range(L):
        for m in range(-ell, ell + 1):
            d = (ell + 1) ** 2 - m * m
            out[ell, m + L, 0] = math.sqrt((2 * ell + 1) * (ell + 1) / d)
            out[ell, m + L, 1] = math.sqrt((ell + 1) * (ell * ell - m * m) / (ell * d)) if ell > 0 else 0.0
    return out


def _phase_powers(z, kmax):
    """[kmax+1, ...] powers z^0..z^kmax by repeated multiplication."""
    out = np.empty((kmax + 1,) + z.shape, dtype=complex)
    out[0] = 1.0
    for k in range(1, kmax + 1):
        out[k] = out[k - 1] * z
    return out


def wigner_D_column(Ra, Rb, ell_max, m_col):
    """D^l_{mp, m_col}(R) for all l<=ell_max, all mp: returns [..., (ell_max+1)^2] indexed LM_index(l,mp,0).

    Entries with l < |m_col| are zero.  The three-term recurrence in l runs for all mp at once (rows that have not
    started yet hold zeros and their table coefficients are zero, so they stay zero until their seed is planted).
    """
    Ra = np.asarray(Ra, dtype=complex)
    Rb = np.asarray(Rb, dtype=complex)
    shape = Ra.shape
    Ra = Ra.reshape(-1)
    Rb = Rb.reshape(-1)
    L = max(ell_max, abs(m_col))
    seed, rec = wigner_tables(L)
    cosb = (Ra.real**2 + Ra.imag**2) - (Rb.real**2 + Rb.imag**2)
    pa = _phase_powers(Ra, 2 * L)
    pb = _phase_powers(Rb, 2 * L)
    m = m_col
    mps = np.arange(-ell_max, ell_max + 1)
    ka, kb = mps + m, m - mps
    ph = np.where((ka >= 0)[:, None], pa[np.abs(ka)], np.conj(pa[np.abs(ka)])) * np.where(
        (kb >= 0)[:, None], pb[np.abs(kb)], np.conj(pb[np.abs(kb)])
    )                                                                       # [n_mp, G]
    l0 = np.maximum(np.abs(mps), abs(m))
    seeds = seed[mps + L, m + L]
    out = np.zeros((Ra.shape[0], (ell_max + 1) ** 2), dtype=complex)
    P = np.zeros((mps.shape[0], Ra.shape[0]))
    Pm1 = np.zeros_like(P)
    for ell in range(abs(m) if abs(m) <= ell_max else ell_max + 1, ell_max + 1):
        start = l0 == ell
        if start.any():
            P[start] = seeds[start][:, None]
        lo, hi = ell_max - ell, ell_max + ell + 1                           # rows with |mp| <= ell
        out[:, ell * ell : (ell + 1) ** 2] = (ph[lo:hi] * P[lo:hi]).T
        if ell < ell_max:
            abc = rec[ell, mps + L, m + L]                                  # [n_mp, 3]; zero rows where l0 > ell
            P, Pm1 = (abc[:, 0:1] * cosb[None, :] - abc[:, 1:2]) * P - abc[:, 2:3] * Pm1, P
    return out.reshape(shape + ((ell_max + 1) ** 2,))


def wigner_D_matrices(Ra, Rb, ell_min, ell_max):
    """Flat D array (sf layout: for l, for mp, for m) for a single rotor or an array of rotors."""
    Ra = np.asarray(Ra, dtype=complex)
    out = np.empty(Ra.shape + (D_total_size(ell_min, ell_max),), dtype=complex)
    for m in range(-ell_max, ell_max + 1):
        col = wigner_D_column(Ra, Rb, ell_max, m)
        for ell in range(max(ell_min, abs(m)), ell_max + 1):
            n = 2 * ell + 1
            base = D_offset(ell, ell_min)
            for mp in range(-ell, ell + 1):
                out[..., base + n * (ell + mp) + (ell + m)] = col[..., LM_index(ell, mp, 0)]
    return out


def SWSH_grid(R, s, ell_max):
    """sY_lm on rotors R[...,4]: (-1)^s sqrt((2l+1)/4pi) D^l_{m,-s}(R), index LM_index(l,m,0), zero for l<|s|.

    Same contract as sf.SWSH_grid (scri/waveform_grid.py:470-471).
    """
    R = np.asarray(R, dtype=float)
    Ra = R[..., 0] + 1j * R[..., 3]
    Rb = R[..., 2] + 1j * R[..., 1]
    out = wigner_D_column(Ra, Rb, ell_max, -s)
    for ell in range(ell_max + 1):
        f = (-1) ** s * math.sqrt((2 * ell + 1) / (4 * math.pi)) if ell >= abs(s) else 0.0
        out[..., ell * ell : (ell + 1) ** 2] *= f
    return out


# ----------------------------------------------------------------------------- analysis quadrature
@lru_cache(maxsize=None)
def clenshaw_curtis_theta_weights(n_theta):
    """q_j with int_0^pi sin(t) g(t) dt ~= sum_j q_j g(pi j/(n_theta-1)).

    This is what Huffenberger & Wandelt's torus extension + weight convolution (spinsfast.map2salm,
    scri/waveform_grid.py:305) reduces to once the ring is folded back onto [0, pi]: the odd part of
    the ring weights cancels between theta and 2pi-theta and the even part is the Clenshaw-Curtis
    rule, the ring's Nyquist frequency entering once.
    """
    N = n_theta - 1
    NG = 2 * N
    j = np.arange(n_theta)
    theta = np.pi * j / N
    W = np.zeros(n_theta)
    for p in range(-(N - 1), N + 1):
        if p % 2 == 0:
            W += 2.0 / (1.0 - p * p) * np.cos(p * theta)
    q = 2.0 * W / NG
    q[0] *= 0.5
    q[-1] *= 0.5
    return q


@lru_cache(maxsize=32)
def analysis_tables(s, ell_min, ell_max, n_theta, n_phi):
    """Tables for map2salm as phi-DFT + theta quadrature.

    Returns
      E[k, mi]  complex [n_phi, 2*ell_max+1] : e^{-i m phi_k}/n_phi, m = mi - ell_max
      Wt[lm, j] real    [n_modes, n_theta]   : 2 pi q_j sY_lm(theta_j, 0)  for (l,m) from ell_min
    so that a_lm = sum_j Wt[lm, j] * sum_k f[j,k] E[k, m+ell_max].
    """
    theta = np.pi * np.arange(n_theta) / (n_theta - 1)
    phi = 2 * np.pi * np.arange(n_phi) / n_phi
    ms = np.arange(-ell_max, ell_max + 1)
    E = np.exp(-1j * phi[:, None] * ms[None, :]) / n_phi
    q = clenshaw_curtis_theta_weights(n_theta)
    R = np.stack([np.cos(theta / 2), np.zeros_like(theta), np.sin(theta / 2), np.zeros_like(theta)], axis=-1)
    Y = SWSH_grid(R, s, ell_max)  # real up to rounding: [n_theta, (L+1)^2]
    Wt = (2 * np.pi * q[None, :] * Y.real.T)[ell_min * ell_min :]
    return E, np.ascontiguousarray(Wt)
