"""Synthetic waveform generators (host numpy), following scri/sample_waveforms.py.

These build the *inputs* of the hot path (tests, bench.py); the only device work is the final
`to_inertial_frame()` of `fake_precessing_waveform(inertial=True)`.
"""
import math
import warnings

import numpy as np
from scipy.interpolate import CubicSpline
from scipy.special import factorial, factorial2

from . import _quaternion as Q
from . import _sf
from .constants import Corotating, Inertial, SpinWeights, h
from .waveform_modes import WaveformModes


def modes_constructor(constructor_statement, data_functor, **kwargs):
    """WaveformModes filled from `data_functor(t, LM)` (scri/sample_waveforms.py:12-67)."""
    t = np.array(kwargs.pop("t", np.linspace(-10.0, 100.0, num=1101)), dtype=float)
    frame = np.array(kwargs.pop("frame", np.empty((0, 4))), dtype=float)
    frameType = int(kwargs.pop("frameType", Inertial))
    dataType = int(kwargs.pop("dataType", h))
    r_is_scaled_out = bool(kwargs.pop("r_is_scaled_out", True))
    m_is_scaled_out = bool(kwargs.pop("m_is_scaled_out", True))
    ell_min = int(kwargs.pop("ell_min", abs(SpinWeights[dataType])))
    ell_max = int(kwargs.pop("ell_max", 8))
    if kwargs:
        import pprint

        warnings.warn(f"\nUnused kwargs passed to this function:\n{pprint.pformat(kwargs, width=1)}")
    data = data_functor(t, _sf.LM_range(ell_min, ell_max))
    return WaveformModes(
        t=t, frame=frame, data=data, history=["# Called from modes_constructor"], frameType=frameType,
        dataType=dataType, r_is_scaled_out=r_is_scaled_out, m_is_scaled_out=m_is_scaled_out,
        constructor_statement=constructor_statement, ell_min=ell_min, ell_max=ell_max,
    )


def constant_waveform(**kwargs):
    """Constant value m - i m in each mode (scri/sample_waveforms.py:70-87)."""

    def data_functor(t, LM):
        data = np.zeros((t.shape[0], LM.shape[0]), dtype=complex)
        for i, m in enumerate(LM[:, 1]):
            data[:, i] = m - 1j * m
        return data

    return modes_constructor(f"constant_waveform(**{kwargs})", data_functor, **kwargs)


def single_mode(ell, m, **kwargs):
    """1 in the (ell, m) slot, 0 elsewhere (scri/sample_waveforms.py:90-110)."""

    def data_functor(t, LM):
        data = np.zeros((t.shape[0], LM.shape[0]), dtype=complex)
        data[:, _sf.LM_index(ell, m, min(LM[:, 0]))] = 1.0 + 0.0j
        return data

    return modes_constructor(f"single_mode({ell}, {m}, **{kwargs})", data_functor, **kwargs)


def single_mode_proportional_to_time(**kwargs):
    """beta * t in one (ell, m) slot of a spin-s field (scri/sample_waveforms.py:256-309)."""
    s = kwargs.pop("s", -2)
    ell = kwargs.pop("ell", abs(s))
    m = kwargs.pop("m", -ell)
    ell_min = kwargs.pop("ell_min", abs(s))
    ell_max = kwargs.pop("ell_max", 8)
    data_type = SpinWeights.index(s)
    t_0 = kwargs.pop("t_0", -20.0)
    t_1 = kwargs.pop("t_1", 20.0)
    dt = kwargs.pop("dt", 1.0 / 10.0)
    t = np.arange(t_0, t_1 + dt, dt)
    n_times = t.size
    beta = kwargs.pop("beta", 1.0)
    data = np.zeros((n_times, _sf.LM_total_size(ell_min, ell_max)), dtype=complex)
    data[:, _sf.LM_index(ell, m, ell_min)] = beta * t
    if kwargs:
        import pprint

        warnings.warn(f"\nUnused kwargs passed to this function:\n{pprint.pformat(kwargs, width=1)}")
    return WaveformModes(
        t=t, data=data, ell_min=ell_min, ell_max=ell_max, frameType=Inertial, dataType=data_type,
        r_is_scaled_out=True, m_is_scaled_out=True,
    )


def smooth_random_waveform(n_times=2048, ell_min=2, ell_max=8, t_0=0.0, t_1=204.7, seed=0, dataType=h, batch=None):
    """Random-mode waveform(s) that are smooth in time: a_lm(t) = c_lm exp(i w_lm t).

    c_lm ~ N(0,1) + i N(0,1), w_lm ~ U(0.05, 0.5): the SURVEY.md 8(d) C3 input (white noise in time makes
    1e-12 spline parity ill-posed).  With `batch=B` returns (t, data[B, n_times, n_modes]).
    """
    rng = np.random.default_rng(seed)
    n = _sf.LM_total_size(ell_min, ell_max)
    t = np.linspace(t_0, t_1, n_times)
    shape = (n,) if batch is None else (batch, 1, n)
    c = rng.normal(size=shape) + 1j * rng.normal(size=shape)
    w = rng.uniform(0.05, 0.5, size=shape)
    if batch is None:
        data = c[None, :] * np.exp(1j * w[None, :] * t[:, None])
        return WaveformModes(
            t=t, data=data, ell_min=ell_min, ell_max=ell_max, frameType=Inertial, dataType=dataType,
            r_is_scaled_out=True, m_is_scaled_out=True,
        )
    data = c * np.exp(1j * w * t[None, :, None])
    return t, data


# ------------------------------------------------------------------------------ utilities.py:11-191
def transition_function(x, x0, x1, y0=0.0, y1=1.0, return_indices=False):
    """C-infinity transition from y0 (x<=x0) to y1 (x>=x1) (scri/utilities.py:11-58)."""
    x = np.asarray(x, dtype=float)
    transition = np.empty_like(x)
    i0 = int(np.searchsorted(x, x0, side="right"))
    i1 = int(np.searchsorted(x, x1, side="left"))
    transition[:i0] = y0
    transition[i1:] = y1
    tau = (x[i0:i1] - x0) / (x1 - x0)
    with np.errstate(over="ignore", divide="ignore"):
        exponent = 1.0 / tau - 1.0 / (1.0 - tau)
        transition[i0:i1] = y0 + (y1 - y0) / (1.0 + np.exp(exponent))
    return (transition, i0, i1) if return_indices else transition


def transition_function_derivative(x, x0, x1, y0=0.0, y1=1.0):
    """d/dx of `transition_function` (scri/utilities.py:62-97)."""
    x = np.asarray(x, dtype=float)
    out = np.zeros_like(x)
    i0 = int(np.searchsorted(x, x0, side="right"))
    i1 = int(np.searchsorted(x, x1, side="left"))
    tau = (x[i0:i1] - x0) / (x1 - x0)
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        exponent = 1.0 / tau - 1.0 / (1.0 - tau)
        e = np.exp(np.minimum(exponent, 700.0))
        d = (y1 - y0) * e * (1.0 / tau**2 + 1.0 / (1.0 - tau) ** 2) / (1.0 + e) ** 2 / (x1 - x0)
    out[i0:i1] = np.where(np.isfinite(d), d, 0.0)
    return out


def indefinite_integral(f, t):
    """Spline antiderivative with zero at t[0] (quaternion.calculus.indefinite_integral)."""
    return CubicSpline(t, f).antiderivative()(t)


def transition_to_constant(f, t, t1, t2):
    """Smoothly freeze `f` to a constant after t2 (scri/utilities.py:155-191)."""
    transition, i1, i2 = transition_function(t, t1, t2, y0=1.0, y1=0.0, return_indices=True)
    transition_dot = transition_function_derivative(t, t1, t2, y0=1.0, y1=0.0)
    f_transitioned = f * transition
    f_transitioned[i1:i2] -= indefinite_integral(f[i1:i2] * transition_dot[i1:i2], t[i1:i2])
    f_transitioned[i2:] = f_transitioned[i2 - 1]
    return f_transitioned


def pn_leading_order_amplitude(ell, m, x, mass_ratio=1.0):
    """Leading-order PN amplitude of r h/M, Blanchet (2014) eq. 330 (scri/sample_waveforms.py:536-593)."""
    if m < 0:
        return (-1) ** ell * np.conjugate(pn_leading_order_amplitude(ell, -m, x, mass_ratio=mass_ratio))
    if mass_ratio < 1.0:
        mass_ratio = 1.0 / mass_ratio
    nu = mass_ratio / (1 + mass_ratio) ** 2
    X1 = mass_ratio / (mass_ratio + 1)
    X2 = 1 / (mass_ratio + 1)

    def sigma(ell):
        return X2 ** (ell - 1) + (-1) ** ell * X1 ** (ell - 1)

    if (ell + m) % 2 == 0:
        amplitude = (
            ((-1) ** ((ell - m + 2) / 2) / (2 ** (ell + 1) * factorial((ell + m) // 2) * factorial((ell - m) // 2) * factorial2(2 * ell - 1)))
            * np.sqrt((5 * (ell + 1) * (ell + 2) * factorial(ell + m) * factorial(ell - m)) / (ell * (ell - 1) * (2 * ell + 1)))
            * sigma(ell) * (1j * m) ** ell * x ** (ell / 2 - 1)
        )
    else:
        amplitude = (
            ((-1) ** ((ell - m - 1) / 2) / (2 ** (ell - 1) * factorial((ell + m - 1) // 2) * factorial((ell - m - 1) // 2) * factorial2(2 * ell + 1)))
            * np.sqrt((5 * (ell + 2) * (2 * ell + 1) * factorial(ell + m) * factorial(ell - m)) / (ell * (ell - 1) * (ell + 1)))
            * sigma(ell + 1) * 1j * (1j * m) ** ell * x ** ((ell - 1) / 2)
        )
    return 8 * np.sqrt(np.pi / 5) * nu * x * amplitude


def _exp_axis(angle, axis):
    """exp(angle * axis / 2) for a unit axis vector: rotor series [n, 4]."""
    angle = np.asarray(angle, dtype=float)
    out = np.zeros(angle.shape + (4,))
    out[..., 0] = np.cos(angle / 2)
    out[..., 1:] = np.sin(angle / 2)[..., None] * np.asarray(axis, dtype=float)
    return out


def fake_precessing_waveform(
    t_0=-20.0, t_1=20_000.0, dt=0.1, ell_max=8, mass_ratio=2.0, precession_opening_angle=np.pi / 6.0,
    precession_opening_angle_dot=None, precession_relative_rate=0.1, precession_nutation_angle=None, inertial=True,
):
    """Strain waveform with realistic precession effects (scri/sample_waveforms.py:383-533).

    With `inertial=True` the corotating-frame modes are rotated back to the inertial frame on the GPU
    (`to_inertial_frame`); `inertial=False` is pure host code.
    """
    if mass_ratio < 1.0:
        mass_ratio = 1.0 / mass_ratio
    s = -2
    ell_min = abs(s)
    nu = mass_ratio / (1 + mass_ratio) ** 2
    t = np.arange(t_0, t_1 + 0.99 * dt, dt)
    t_merger = t_1 - 100.0
    i_merger = np.argmin(abs(t - t_merger))
    if i_merger < 20:
        raise ValueError(f"Insufficient space between initial time (t={t_merger}) and merger (t={t_0}).")
    n_times = t.size
    data = np.zeros((n_times, _sf.LM_total_size(ell_min, ell_max)), dtype=complex)

    tau = nu * (t_merger - t) / 5
    with np.errstate(invalid="ignore", divide="ignore"):
        phi = -4 * tau ** (5 / 8)
        omega = (nu / 2) * tau ** (-3 / 8)

    omega_transition_width = 5.0
    i1 = np.argmin(np.abs(omega[~np.isnan(omega)] - 0.25))
    i0 = np.argmin(np.abs(t - (t[i1] - omega_transition_width)))
    transition = transition_function(t, t[i0], t[i1])
    omega[:i1] = omega[:i1] * (1 - transition[:i1]) + 0.25 * transition[:i1]
    omega[i1:] = 0.25
    phi[i0:] = phi[i0] + indefinite_integral(omega[i0:], t[i0:])

    ringdown_transition_width = 20
    i0 = np.argmin(np.abs(t - t_merger))
    i1 = np.argmin(np.abs(t - (t[i0] + ringdown_transition_width)))
    t0, t1 = t[i0], t[i1]
    transition = transition_function(t, t0, t1)
    ringdown = np.ones_like(t)
    ringdown[i0:] = ringdown[i0:] * (1 - transition[i0:]) + 2.25 * np.exp(-(t[i0:] - t_merger) / 11.5) * transition[i0:]

    if precession_opening_angle_dot is None:
        precession_opening_angle_dot = 2.0 * precession_opening_angle / (t[i1] - t[0])
    if precession_nutation_angle is None:
        precession_nutation_angle = precession_opening_angle / 10.0
    ex, ez = [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]
    R_orbital = _exp_axis(phi, ez)
    R_opening = _exp_axis(transition_to_constant(precession_opening_angle + precession_opening_angle_dot * t, t, t0, t1), ex)
    R_precession = _exp_axis(transition_to_constant(phi / precession_relative_rate, t, t0, t1), ez)
    R_nutation = _exp_axis(precession_nutation_angle * transition, ex)
    frame = R_orbital
    for fac in (R_nutation, Q.qconj(R_orbital), R_precession, R_opening, Q.qconj(R_precession), R_orbital):
        frame = Q.qmul(frame, fac)
    frame = Q.qmul(Q.qconj(Q.qsqrt(frame[0])), frame)

    x = omega ** (2 / 3)
    modulation = transition_function(t, t[i0], t[i1], 1, 0) * np.cos(phi) / 40.0
    for ell in range(ell_min, ell_max + 1):
        for m in range(-ell, ell + 1):
            data[:, _sf.LM_index(ell, m, ell_min)] = pn_leading_order_amplitude(ell, m, x, mass_ratio=mass_ratio) * (1 + np.sign(m) * modulation)
    data *= ringdown[:, np.newaxis]

    h_corot = WaveformModes(
        t=t, frame=frame, data=data, ell_min=ell_min, ell_max=ell_max, frameType=Corotating, dataType=h,
        r_is_scaled_out=True, m_is_scaled_out=True,
    )
    if inertial:
        return h_corot.to_inertial_frame()
    return h_corot
