"""Restatement of the reference's frame logic and codec pre-stages on float rotor arrays.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows scri/mode_calculations.py:435-490 (corotating_frame),
scri/rotations.py:14-48 (to_coprecessing_frame), :51-103 (to_corotating_frame), scri/utilities.py:11-58
(transition_function), scri/waveform_modes.py:457-476 (truncate), :658-703 (conjugate pairs),
scri/waveform_base.py:553-575 (max_norm_time), :949-967 (interpolate).  Pinned by tests/golden/reference_frames.npz and
reference_codec.npz, made by running the reference itself (tests/golden/make_reference_vectors.py).
"""
import math

import numpy as np

from . import quat, quat_series
from . import scri_ref as R


def max_norm_time(W, skip_fraction_of_data=4):
    """waveform_base.py:553-575"""
    n = R.norm(W)
    if skip_fraction_of_data in (0, 1):
        return W.t[int(np.argmax(n))]
    k = W.n_times // skip_fraction_of_data
    return W.t[int(np.argmax(n[k:])) + k]


def _slice(W, i1, i2):
    return R.Modes(t=W.t[i1:i2], data=W.data[i1:i2], ell_min=W.ell_min, ell_max=W.ell_max, dataType=W.dataType)


def corotating_frame(W, R0=None, tolerance=1e-12, z_alignment_region=None, return_omega=False):
    """mode_calculations.py:435-490"""
    omega = R.angular_velocity(W)
    frame = quat_series.integrate_angular_velocity(W.t, omega, R0, tolerance)
    if z_alignment_region is None:
        correction_rotor = quat.one
    else:
        initial_time = W.t[0]
        inspiral_time = max_norm_time(W) - initial_time
        t1 = initial_time + z_alignment_region[0] * inspiral_time
        t2 = initial_time + z_alignment_region[1] * inspiral_time
        i1 = int(np.argmin(np.abs(W.t - t1)))
        i2 = int(np.argmin(np.abs(W.t - t2)))
        Rr = frame[i1:i2]
        i1m = max(0, i1 - 10)
        RoughDirection = omega[i1m + 10]
        Vhat = R.LLDominantEigenvector(_slice(W, i1, i2), RoughDirection=RoughDirection, RoughDirectionIndex=0)
        # (Ri.conjugate() * quaternion(*Vhati) * Ri).vec
        Vq = np.concatenate([np.zeros((Vhat.shape[0], 1)), Vhat], axis=1)
        Vhat_corot = quat.mul(quat.mul(quat.conj(Rr), Vq), Rr)[:, 1:]
        mean = np.mean(Vhat_corot, axis=0)
        mq = quat.normalized(np.concatenate([[0.0], mean]))
        correction_rotor = quat.inverse(quat.sqrt(quat.mul(-quat.z, mq)))
    frame = quat.mul(frame, correction_rotor)
    frame = frame / quat.absq(frame)[:, None]
    return (frame, omega) if return_omega else frame


def to_corotating_frame(W, R0=None, tolerance=1e-12, z_alignment_region=None, truncate_log_frame=False):
    """rotations.py:51-103; returns (W', omega, log_frame or None)."""
    frame, omega = corotating_frame(W, R0=R0, tolerance=tolerance, z_alignment_region=z_alignment_region, return_omega=True)
    log_frame = None
    if truncate_log_frame:
        log_frame = quat.log(frame)
        power_of_2 = 2 ** int(-np.floor(np.log2(2 * tolerance)))
        log_frame = np.round(log_frame * power_of_2) / power_of_2
        frame = quat.exp(log_frame)
    out = R.rotate_decomposition_basis(W.copy(), frame)
    return out, omega, log_frame


def transition_function(x, x0, x1, y0=0.0, y1=1.0):
    """utilities.py:11-58"""
    maxexp = np.finfo(float).maxexp * np.log(2) * 0.99
    out = np.empty_like(x)
    for i, xi in enumerate(x):
        if xi <= x0:
            out[i] = y0
        elif xi < x1:
            tau = (xi - x0) / (x1 - x0)
            e = 1.0 / tau - 1.0 / (1.0 - tau)
            out[i] = y0 if e >= maxexp else y0 + (y1 - y0) / (1.0 + math.exp(e))
        else:
            out[i] = y1
    return out


def coprecessing_frame(W, RoughDirection=np.array([0.0, 0.0, 1.0]), RoughDirectionIndex=None, transition_times=None):
    """The rotor series of rotations.py:14-48."""
    if RoughDirectionIndex is None:
        RoughDirectionIndex = W.n_times // 8
    dpa = R.LLDominantEigenvector(W, RoughDirection=RoughDirection, RoughDirectionIndex=RoughDirectionIndex)
    dq = quat.normalized(np.concatenate([np.zeros((dpa.shape[0], 1)), dpa], axis=1))
    Rf = quat.sqrt(quat.mul(-dq, quat.z))
    Rf = quat_series.minimal_rotation(Rf, W.t, iterations=3)
    if transition_times is not None:
        i0 = int(np.argmin(np.abs(W.t - transition_times[0])))
        i1 = int(np.argmin(np.abs(W.t - transition_times[1])))
        transition = transition_function(W.t[i0:], W.t[i0], W.t[i1], y0=1.0, y1=0.0)
        omega = quat_series.angular_velocity(Rf[i0:], W.t[i0:]) * transition[:, np.newaxis]
        slowing = quat_series.integrate_angular_velocity(W.t[i0:], omega, Rf[i0])
        Rf = np.concatenate((Rf[:i0], slowing))
    return Rf


def to_coprecessing_frame(W, **kwargs):
    frame = coprecessing_frame(W, **kwargs)
    return R.rotate_decomposition_basis(W.copy(), frame), frame


def interpolate(W, tprime, frame=None):
    """waveform_base.py:949-967: CubicSpline for the data, squad for the frame."""
    data = R.interpolate_data(W, tprime)
    fr = quat_series.squad(frame, W.t, tprime) if frame is not None and len(frame) == W.n_times else frame
    return data, fr


def convert_to_conjugate_pairs(data, ell_min, ell_max):
    """waveform_modes.py:658-686"""
    data = data.copy()
    for ell in range(ell_min, ell_max + 1):
        for m in range(1, ell + 1):
            ip = ell * (ell + 1) - ell_min**2 + m
            im = ell * (ell + 1) - ell_min**2 - m
            p, q = data[..., ip].copy(), data[..., im].copy()
            data[..., ip] = (p + np.conjugate(q)) / np.sqrt(2)
            data[..., im] = (p - np.conjugate(q)) / np.sqrt(2)
    return data


def convert_from_conjugate_pairs(data, ell_min, ell_max):
    """waveform_modes.py:688-703"""
    data = data.copy()
    for ell in range(ell_min, ell_max + 1):
        for m in range(1, ell + 1):
            ip = ell * (ell + 1) - ell_min**2 + m
            im = ell * (ell + 1) - ell_min**2 - m
            p, q = data[..., ip].copy(), data[..., im].copy()
            data[..., ip] = (p + q) / np.sqrt(2)
            data[..., im] = np.conjugate(p - q) / np.sqrt(2)
    return data


def truncate(data, tol=1e-10):
    """waveform_modes.py:457-476"""
    data = data.copy()
    if tol != 0.0:
        tol_per_mode = tol / np.sqrt(data.shape[1])
        absolute_tolerance = np.linalg.norm(data, axis=1) * tol_per_mode
        power_of_2 = (2.0 ** np.floor(-np.log2(absolute_tolerance)))[:, np.newaxis]
        data *= power_of_2
        np.round(data, out=data)
        data /= power_of_2
    return data
