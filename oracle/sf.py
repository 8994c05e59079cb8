"""Restatement of the `spherical_functions` (sf) routines on the scri hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  `spherical-functions>=2022.4`
(reference pyproject.toml) is not vendored under /root/reference.  Call sites followed:
scri/waveform_grid.py:49,60-86,470-471,488,497,508 ; scri/rotations.py:299,327,359,381-389 ;
scri/waveform_modes.py:139,437,521-528 ; scri/mode_calculations.py:9-10 ; scri/flux.py:7,225,267-284,371-375 ;
scri/sample_waveforms.py:353-364.

Formulas: SURVEY.md Appendix A.1/A.3 (published definitions in the sf documentation,
http://moble.github.io/spherical_functions/).  The Wigner-D sum is evaluated in extended
precision (np.longdouble) up to l = 16 and through its Jacobi-polynomial closed form above (the
alternating sum cancels there), so the oracle stays at the 1e-14 level up to l = 64 (config 4 works at
l = 32 / 64); agreement with sf is expected at the 1e-15..1e-14 level, not bit-for-bit.
"""
import math
from functools import lru_cache

import numpy as np


# ----------------------------------------------------------------------------- indexing (A.1)
def LM_index(ell, m, ell_min):
    return ell * (ell + 1) - ell_min**2 + m


def LM_total_size(ell_min, ell_max):
    return ell_max * (ell_max + 2) - ell_min**2 + 1


def LM_range(ell_min, ell_max):
    return np.array([[ell, m] for ell in range(ell_min, ell_max + 1) for m in range(-ell, ell + 1)], dtype=np.int64)


def linear_matrix_offset(ell, ell_min):
    """sf._linear_matrix_offset: start of the ell block in the flat D array."""
    return ((4 * ell**2 - 1) * ell - (4 * ell_min**2 - 1) * ell_min) // 3


def total_size_D_matrices(ell_min, ell_max):
    return sum((2 * ell + 1) ** 2 for ell in range(ell_min, ell_max + 1))


# ----------------------------------------------------------------------------- Wigner D (A.3)
_EXPLICIT_SUM_ELL_MAX = 16


def _wigner_D_element_sum(Ra, Rb, ell, mp, m):
    """D^ell_{mp,m}(R) by the explicit sum, in long double; Ra, Rb complex arrays (any shape)."""
    Ra = np.asarray(Ra, dtype=np.clongdouble)
    Rb = np.asarray(Rb, dtype=np.clongdouble)
    ra2 = (Ra * np.conj(Ra)).real
    rb2 = (Rb * np.conj(Rb)).real
    pref = math.sqrt(
        math.factorial(ell + m) * math.factorial(ell - m) / (math.factorial(ell + mp) * math.factorial(ell - mp))
    )
    # exact in long double for the ells we use
    pref = np.longdouble(math.factorial(ell + m) * math.factorial(ell - m)) / np.longdouble(
        math.factorial(ell + mp) * math.factorial(ell - mp)
    )
    pref = np.sqrt(pref)
    rho_min = max(0, mp - m)
    rho_max = min(ell + mp, ell - m)
    # term = C C (-1)^rho Ra^(l+mp-rho) conj(Ra)^(l-rho-m) Rb^(rho-mp+m) conj(Rb)^rho
    #      = PhA(mp+m) PhB(m-mp) * ra2^(...) rb2^(...)
    ka = mp + m
    kb = m - mp
    # phases without division: Ra^ka if ka>=0 else conj(Ra)^|ka| ; remaining modulus powers are even
    total = np.zeros(Ra.shape, dtype=np.longdouble)
    for rho in range(rho_min, rho_max + 1):
        c = math.comb(ell + mp, rho) * math.comb(ell - mp, ell - rho - m)
        ea = ell + mp - rho  # power of Ra
        eca = ell - rho - m  # power of conj(Ra)
        eb = rho - mp + m  # power of Rb
        ecb = rho  # power of conj(Rb)
        pa = min(ea, eca)  # |Ra|^2 power
        pb = min(eb, ecb)
        assert ea - pa == max(ka, 0) and eca - pa == max(-ka, 0)
        assert eb - pb == max(kb, 0) and ecb - pb == max(-kb, 0)
        total = total + np.longdouble((-1) ** rho * c) * ra2**pa * rb2**pb
    phase = (Ra ** max(ka, 0)) * (np.conj(Ra) ** max(-ka, 0)) * (Rb ** max(kb, 0)) * (np.conj(Rb) ** max(-kb, 0))
    return pref * total * phase


def _wigner_D_element_ld(Ra, Rb, ell, mp, m):
    """D^ell_{mp,m}(R); Ra, Rb complex arrays (any shape), |Ra|^2 + |Rb|^2 = 1.

    The explicit sum of SURVEY.md A.3,
        D = sqrt((l+m)!(l-m)!/((l+mp)!(l-mp)!)) sum_rho (-1)^rho C(l+mp,rho) C(l-mp,l-rho-m)
                Ra^(l+mp-rho) conj(Ra)^(l-rho-m) Rb^(rho-mp+m) conj(Rb)^rho,
    factors into the phases Ra^(mp+m) Rb^(m-mp) (conjugates for negative powers) and a real alternating sum in
    |Ra|^2, |Rb|^2.  That sum cancels catastrophically for l >~ 25 (1e-11 at l = 32 even in long double), so it is
    evaluated in its closed form instead: it is Wigner's d^l_{mp,m}(beta) / (|Ra|^|mp+m| |Rb|^|m-mp|) with
    cos(beta) = |Ra|^2 - |Rb|^2, i.e. a Jacobi polynomial (stable three-term recurrence, scipy.special.eval_jacobi):
        sum = (-1)^max(mp-m,0) sqrt(C(2l-k, k+a) / C(k+b, b)) P_k^(a,b)(cos beta),
        k = min(l+m, l-m, l+mp, l-mp), a = |mp-m|, b = |mp+m|.
    Checked against a 60-digit evaluation of the sum up to l = 64 (tests/test_oracle.py): 4e-14.  Up to
    l = _EXPLICIT_SUM_ELL_MAX the long-double sum itself is used (`_wigner_D_element_sum`, < 1e-15 there)."""
    from scipy.special import eval_jacobi

    if ell <= _EXPLICIT_SUM_ELL_MAX:
        return _wigner_D_element_sum(Ra, Rb, ell, mp, m)
    Ra = np.asarray(Ra, dtype=complex)
    Rb = np.asarray(Rb, dtype=complex)
    ra2 = (Ra * np.conj(Ra)).real
    rb2 = (Rb * np.conj(Rb)).real
    ka = mp + m
    kb = m - mp
    k = min(ell + m, ell - m, ell + mp, ell - mp)
    a, b = abs(kb), abs(ka)
    coef = math.sqrt(math.comb(2 * ell - k, k + a) / math.comb(k + b, b))
    total = ((-1) ** max(mp - m, 0) * coef) * eval_jacobi(k, a, b, ra2 - rb2)
    phase = (Ra ** max(ka, 0)) * (np.conj(Ra) ** max(-ka, 0)) * (Rb ** max(kb, 0)) * (np.conj(Rb) ** max(-kb, 0))
    return total * phase


def Wigner_D_matrices(Ra, Rb, ell_min, ell_max):
    """Flat D array in sf's layout (A.1): for ell, for mp, for m. Leading dims = Ra.shape."""
    Ra = np.asarray(Ra, dtype=complex)
    Rb = np.asarray(Rb, dtype=complex)
    out = np.empty(Ra.shape + (total_size_D_matrices(ell_min, ell_max),), dtype=complex)
    i = 0
    for ell in range(ell_min, ell_max + 1):
        for mp in range(-ell, ell + 1):
            for m in range(-ell, ell + 1):
                out[..., i] = _wigner_D_element_ld(Ra, Rb, ell, mp, m).astype(complex)
                i += 1
    return out


def Wigner_D_element(Ra, Rb, ell, mp, m):
    return _wigner_D_element_ld(Ra, Rb, ell, mp, m).astype(complex)


_swsh_cache = {}


def SWSH_grid(R, s, ell_max):
    """Memoised front end of `_SWSH_grid` (the long-double evaluation is slow; sf's own is numba-fast)."""
    R = np.ascontiguousarray(R, dtype=float)
    key = (R.tobytes(), R.shape, s, ell_max)
    if key not in _swsh_cache:
        if len(_swsh_cache) > 64:
            _swsh_cache.clear()
        _swsh_cache[key] = _SWSH_grid(R, s, ell_max)
    return _swsh_cache[key].copy()


def _SWSH_grid(R, s, ell_max):
    """sf.SWSH_grid: Y[..., LM_index(l,m,0)] = (-1)^s sqrt((2l+1)/4pi) D^l_{m,-s}(R); zeros for l<|s|.

    R: float array [..., 4].  (scri/waveform_grid.py:470-471)
    """
    R = np.asarray(R, dtype=float)
    Ra = R[..., 0] + 1j * R[..., 3]
    Rb = R[..., 2] + 1j * R[..., 1]
    out = np.zeros(R.shape[:-1] + ((ell_max + 1) ** 2,), dtype=complex)
    for ell in range(abs(s), ell_max + 1):
        f = (-1) ** s * math.sqrt((2 * ell + 1) / (4 * math.pi))
        for m in range(-ell, ell + 1):
            out[..., LM_index(ell, m, 0)] = f * _wigner_D_element_ld(Ra, Rb, ell, m, -s).astype(complex)
    return out


def SWSH(R, s, ell, m):
    R = np.asarray(R, dtype=float)
    Ra = R[..., 0] + 1j * R[..., 3]
    Rb = R[..., 2] + 1j * R[..., 1]
    if ell < abs(s):
        return np.zeros(R.shape[:-1], dtype=complex)
    f = (-1) ** s * math.sqrt((2 * ell + 1) / (4 * math.pi))
    return f * _wigner_D_element_ld(Ra, Rb, ell, m, -s).astype(complex)


def wigner_d_small(beta, ell, mp, m):
    """Real d^ell_{mp,m}(beta) in sf's convention: D(alpha,beta,gamma)=e^{i mp alpha} d e^{i m gamma}...
    evaluated as D on the rotor exp(beta y/2): Ra=cos(beta/2), Rb=sin(beta/2)."""
    beta = np.asarray(beta, dtype=np.longdouble)
    return _wigner_D_element_ld(np.cos(beta / 2), np.sin(beta / 2), ell, mp, m).real


# ----------------------------------------------------------------------------- ladder / CG / 3j
def ladder_operator_coefficient(ell, m):
    """sqrt((l-m)(l+m+1))  (scri/flux.py:371-372 uses the same expression)"""
    return math.sqrt((ell - m) * (ell + m + 1))


def Wigner3j(j1, j2, j3, m1, m2, m3):
    """Standard Wigner 3-j symbol (Racah formula, exact rational arithmetic under the square root)."""
    return _wigner3j(int(j1), int(j2), int(j3), int(m1), int(m2), int(m3))


@lru_cache(maxsize=None)
def _wigner3j(j1, j2, j3, m1, m2, m3):
    from fractions import Fraction

    if m1 + m2 + m3 != 0:
        return 0.0
    if abs(m1) > j1 or abs(m2) > j2 or abs(m3) > j3:
        return 0.0
    if j3 < abs(j1 - j2) or j3 > j1 + j2:
        return 0.0
    f = math.factorial
    tri = Fraction(f(j1 + j2 - j3) * f(j1 - j2 + j3) * f(-j1 + j2 + j3), f(j1 + j2 + j3 + 1))
    pre = tri * f(j1 + m1) * f(j1 - m1) * f(j2 + m2) * f(j2 - m2) * f(j3 + m3) * f(j3 - m3)
    kmin = max(0, j2 - j3 - m1, j1 - j3 + m2)
    kmax = min(j1 + j2 - j3, j1 - m1, j2 + m2)
    s = Fraction(0)
    for k in range(kmin, kmax + 1):
        s += Fraction(
            (-1) ** k,
            f(k) * f(j1 + j2 - j3 - k) * f(j1 - m1 - k) * f(j2 + m2 - k) * f(j3 - j2 + m1 + k) * f(j3 - j1 - m2 + k),
        )
    sign = -1 if (j1 - j2 - m3) % 2 else 1
    # pre * s^2 is an exact rational; take one square root at the end
    val2 = pre * s * s
    return sign * (1 if s >= 0 else -1) * math.sqrt(val2.numerator) / math.sqrt(val2.denominator)


def clebsch_gordan(j1, m1, j2, m2, j3, m3):
    """<j1 m1 j2 m2 | j3 m3>  (sf argument order; scri/flux.py:225)"""
    if m1 + m2 != m3:
        return 0.0
    return (-1) ** (j1 - j2 + m3) * math.sqrt(2 * j3 + 1) * Wigner3j(j1, j2, j3, m1, m2, -m3)


# ----------------------------------------------------------------------------- eth operators (A.3)
def _ell_of_modes(n_modes, ell_min):
    ells = []
    ell = ell_min
    while len(ells) < n_modes:
        ells.extend([ell] * (2 * ell + 1))
        ell += 1
    return np.array(ells[:n_modes], dtype=float)


def eth_GHP(modes, spin_weight, ell_min=0):
    """Multiply by sqrt((l-s)(l+s+1)/2)  (scri/waveform_grid.py:497,508; waveform_modes.py:521-528)"""
    modes = np.asarray(modes)
    ell = _ell_of_modes(modes.shape[-1], ell_min)
    s = spin_weight
    fac = np.where(ell >= abs(s), np.sqrt(np.maximum((ell - s) * (ell + s + 1), 0.0) / 2.0), 0.0)
    return modes * fac


def ethbar_GHP(modes, spin_weight, ell_min=0):
    """Multiply by -sqrt((l+s)(l-s+1)/2)  (scri/waveform_grid.py:488)"""
    modes = np.asarray(modes)
    ell = _ell_of_modes(modes.shape[-1], ell_min)
    s = spin_weight
    fac = np.where(ell >= abs(s), -np.sqrt(np.maximum((ell + s) * (ell - s + 1), 0.0) / 2.0), 0.0)
    return modes * fac


# ----------------------------------------------------------------------------- ell=0,1 conversions
def constant_as_ell_0_mode(c):
    return c * math.sqrt(4 * math.pi)


def constant_from_ell_0_mode(mode):
    return mode / math.sqrt(4 * math.pi)


def vector_as_ell_1_modes(v):
    v = np.asarray(v)
    return np.array(
        [
            math.sqrt(2 * math.pi / 3) * (v[0] + 1j * v[1]),
            math.sqrt(4 * math.pi / 3) * v[2],
            math.sqrt(2 * math.pi / 3) * (-v[0] + 1j * v[1]),
        ],
        dtype=complex,
    )


def vector_from_ell_1_modes(modes):
    a0, a1, a2 = modes
    return np.array(
        [
            (a0 - a2) / (2 * math.sqrt(2 * math.pi / 3)),
            (a0 + a2) / (2j * math.sqrt(2 * math.pi / 3)),
            a1 / math.sqrt(4 * math.pi / 3),
        ]
    )


# ----------------------------------------------------------------------------- Modes.multiply (3j product)
def modes_multiply(a, s1, ell1_max, b, s2, ell2_max, ell_out):
    """Exact mode weights of the product of two spin-weighted functions, truncated at ell_out: the sum over Wigner 3j
    symbols that spherical_functions' Modes.multiply evaluates (called at scri/asymptotic_bondi_data/bms_charges.py:40-187):
      (fg)_{lm} = sum f_{l1 m1} g_{l2 m2} (-1)^{m+s} sqrt((2l1+1)(2l2+1)(2l+1)/4pi)
                      (l1 l2 l; m1 m2 -m) (l1 l2 l; -s1 -s2 s),     s = s1 + s2, m = m1 + m2,
    modes of both factors and of the result stored from ell = 0.  Pinned by tests/test_oracle.py against the grid
    product of the pinned spinsfast restatement (the two are independent computations of the same coefficients)."""
    s = s1 + s2
    out = np.zeros(a.shape[:-1] + ((ell_out + 1) ** 2,), dtype=complex)
    for l1 in range(abs(s1), ell1_max + 1):
        for l2 in range(abs(s2), ell2_max + 1):
            for l in range(max(abs(l1 - l2), abs(s)), min(l1 + l2, ell_out) + 1):
                w_s = Wigner3j(l1, l2, l, -s1, -s2, s)
                if w_s == 0.0:
                    continue
                pref = math.sqrt((2 * l1 + 1) * (2 * l2 + 1) * (2 * l + 1) / (4 * math.pi)) * w_s
                for m1 in range(-l1, l1 + 1):
                    for m2 in range(max(-l2, -l - m1), min(l2, l - m1) + 1):
                        m = m1 + m2
                        w = Wigner3j(l1, l2, l, m1, m2, -m)
                        if w != 0.0:
                            out[..., l * (l + 1) + m] += ((-1) ** ((m + s) % 2) * pref * w) * a[..., l1 * (l1 + 1) + m1] * b[..., l2 * (l2 + 1) + m2]
    return out
