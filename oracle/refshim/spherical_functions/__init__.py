"""Stand-in for `spherical_functions` so that the UNMODIFIED reference (/root/reference/scri) can be imported here.

TEST INFRASTRUCTURE (see oracle/__init__.py).  spherical-functions (>=2022.4 in the reference's pyproject.toml) is not
vendored in /root/reference and not installable in this image.  This module restates the part of its API that scri calls;
the functions scri calls from inside its own numba loops (`_Wigner_D_matrices`, `_linear_matrix_offset`,
`ladder_operator_coefficient`, `LM_index`) are numba functions here too, so the reference's jitted loops
(scri/rotations.py:346-392, scri/mode_calculations.py:14-363, scri/flux.py:40-78) run UNCHANGED on top of them.
The Wigner-D elements follow the published closed form evaluated as sf does, from the polar forms of
(Ra, Rb) with one real alternating sum per element; bit-level agreement with sf's rounding is not claimed.
"""
import math

import numba
import numpy as np

from oracle import sf as _sf

__version__ = "2022.4+oracle.shim"

njit = numba.njit


@njit
def LM_index(ell, m, ell_min):
    return ell * (ell + 1) - ell_min**2 + m


@njit
def LM_total_size(ell_min, ell_max):
    return ell_max * (ell_max + 2) - ell_min**2 + 1


def LM_range(ell_min, ell_max):
    return np.array([[ell, m] for ell in range(ell_min, ell_max + 1) for m in range(-ell, ell + 1)], dtype=int).reshape(-1, 2)


@njit
def _linear_matrix_offset(ell, ell_min):
    return ((4 * ell**2 - 1) * ell - (4 * ell_min**2 - 1) * ell_min) // 3


@njit
def _linear_matrix_index(ell, mp, m):
    return (ell + mp) * (2 * ell + 1) + ell + m


@njit
def _total_size_D_matrices(ell_min, ell_max):
    return _linear_matrix_offset(ell_max + 1, ell_min)


@njit
def ladder_operator_coefficient(ell, m):
    return math.sqrt((ell - m) * (ell + m + 1.0)) if abs(m) <= ell else 0.0


_BINOM_N = 130
_binom = np.zeros((_BINOM_N, _BINOM_N))
for _n in range(_BINOM_N):
    _binom[_n, 0] = 1.0
    for _k in range(1, _n + 1):
        _binom[_n, _k] = _binom[_n - 1, _k - 1] + _binom[_n - 1, _k]


@njit
def _Wigner_D_matrices(Ra, Rb, ell_min, ell_max, elements):
    """elements[offset(ell) + (2 ell + 1)(ell + mp) + ell + m] = D^ell_{mp,m}(R) for the rotor with spinor parts
    Ra = w + i z, Rb = y + i x."""
    ra = abs(Ra)
    rb = abs(Rb)
    phia = math.atan2(Ra.imag, Ra.real)
    phib = math.atan2(Rb.imag, Rb.real)
    for ell in range(ell_min, ell_max + 1):
        i_ell = _linear_matrix_offset(ell, ell_min)
        for mp in range(-ell, ell + 1):
            for m in range(-ell, ell + 1):
                rho_min = max(0, mp - m)
                rho_max = min(ell + mp, ell - m)
                # sqrt[(l+m)!(l-m)!/((l+mp)!(l-mp)!)] C(l+mp, rho) C(l-mp, l-rho-m)
                #   = sqrt[C(2l, l+m)/C(2l, l+mp)] C(l+mp, rho) C(l-mp, l-rho-m)   (ratios of binomials stay small)
                pref = math.sqrt(_binom[2 * ell, ell + mp] / _binom[2 * ell, ell + m])
                total = 0.0
                for rho in range(rho_min, rho_max + 1):
                    term = _binom[ell + mp, rho] * _binom[ell - mp, ell - rho - m]
                    ea = 2 * ell + mp - m - 2 * rho
                    eb = 2 * rho + m - mp
                    term *= ra**ea * rb**eb
                    if rho % 2 == 1:
                        term = -term
                    total += term
                # C(2l,l+mp)/C(2l,l+m) = (l+m)!(l-m)!/((l+mp)!(l-mp)!)
                ang = phia * (mp + m) + phib * (m - mp)
                elements[i_ell + (2 * ell + 1) * (ell + mp) + ell + m] = (pref * total) * complex(math.cos(ang), math.sin(ang))


LMpM_total_size = _total_size_D_matrices


class _WignerDNamespace:
    _total_size_D_matrices = staticmethod(_total_size_D_matrices)
    _linear_matrix_offset = staticmethod(_linear_matrix_offset)
    _linear_matrix_index = staticmethod(_linear_matrix_index)


WignerD = _WignerDNamespace


def Wigner_D_matrices(R, ell_min, ell_max):
    import quaternion

    Rf = quaternion.as_float_array(R)
    flat = Rf.reshape(-1, 4)
    out = np.empty((flat.shape[0], _total_size_D_matrices(ell_min, ell_max)), dtype=complex)
    for i, r in enumerate(flat):
        _Wigner_D_matrices(complex(r[0], r[3]), complex(r[2], r[1]), ell_min, ell_max, out[i])
    return out.reshape(Rf.shape[:-1] + (-1,))


def Wigner_D_element(R, ell, mp, m):
    D = Wigner_D_matrices(R, ell, ell)
    return D[..., _linear_matrix_index(ell, mp, m)]


def SWSH_grid(R, s, ell_max):
    """Y[..., LM_index(l, m, 0)] = sYlm(R) = (-1)^s sqrt((2l+1)/4pi) D^l_{m,-s}(R); zero below l = |s|.  Evaluated by
    oracle/sf.py (extended precision / Jacobi form: 1e-14 up to l = 64; the ABD tests of the reference work at l = 32)."""
    import quaternion

    return _sf._SWSH_grid(quaternion.as_float_array(R), s, ell_max)


def SWSH(R, s, indices):
    indices = np.asarray(indices)
    ell_max = int(indices[..., 0].max())
    Y = SWSH_grid(R, s, ell_max)
    idx = indices[..., 0] * (indices[..., 0] + 1) + indices[..., 1]
    return Y[..., idx]


def Wigner3j(j1, j2, j3, m1, m2, m3):
    return _sf.Wigner3j(j1, j2, j3, m1, m2, m3)


def clebsch_gordan(j1, m1, j2, m2, j3, m3):
    return _sf.clebsch_gordan(j1, m1, j2, m2, j3, m3)


def eth_GHP(modes, spin_weight, ell_min=0):
    s = spin_weight
    return _sf.eth_GHP(modes, s, ell_min)


def ethbar_GHP(modes, spin_weight, ell_min=0):
    s = spin_weight
    return _sf.ethbar_GHP(modes, s, ell_min)


def eth_NP(modes, spin_weight, ell_min=0):
    s = spin_weight
    return math.sqrt(2) * _sf.eth_GHP(modes, s, ell_min)


def ethbar_NP(modes, spin_weight, ell_min=0):
    s = spin_weight
    return math.sqrt(2) * _sf.ethbar_GHP(modes, s, ell_min)


constant_as_ell_0_mode = _sf.constant_as_ell_0_mode
constant_from_ell_0_mode = _sf.constant_from_ell_0_mode
vector_as_ell_1_modes = _sf.vector_as_ell_1_modes
vector_from_ell_1_modes = _sf.vector_from_ell_1_modes


def theta_phi(n_theta, n_phi):
    """[n_theta, n_phi, 2] of (theta_j, phi_k): theta inclusive of both poles, phi half open (spinsfast's grid)."""
    return np.array(
        [[[theta, phi] for phi in np.linspace(0.0, 2 * np.pi, num=n_phi, endpoint=False)]
         for theta in np.linspace(0.0, np.pi, num=n_theta, endpoint=True)]
    )


from .SWSH_modes import Modes  # noqa: E402
from .SWSH_grids import Grid  # noqa: E402
from . import SWSH_modes, SWSH_grids  # noqa: E402,F401
