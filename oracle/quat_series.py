"""Rotor time series: squad, integrate_angular_velocity, minimal_rotation, angular_velocity on float arrays [N, 4].

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates numpy-quaternion's `quaternion_time_series` routines (third party,
>=2024.0.2, not vendored in /root/reference) from their published algorithms - Shoemake's squad with the unequal-step
control points, Boyle's "integration of angular velocity" (dR/dt = (1/2) omega R with a cubic-spline omega and an adaptive
Runge-Kutta), and the minimal-rotation condition through spline derivative + antiderivative.  The reference calls them at
scri/mode_calculations.py:464-467, scri/rotations.py:38-43, scri/waveform_base.py:634-646,962.  oracle/refshim/quaternion
wraps these same functions, so the goldens made by running the real scri and this restatement share that arithmetic;
step selection / rounding of the real package stay unpinned (reference's own bar for the corotating frame: 1e-10).
"""
import numpy as np
from scipy.interpolate import CubicSpline, InterpolatedUnivariateSpline

from . import quat as _q


def _spline_derivative(f, t):
    f = np.asarray(f, dtype=float)
    cols = f.reshape(t.size, -1)
    out = np.empty_like(cols)
    for j in range(cols.shape[1]):
        out[:, j] = InterpolatedUnivariateSpline(t, cols[:, j], k=3).derivative(1)(t)
    return out.reshape(f.shape)


def _spline_antiderivative(f, t):
    f = np.asarray(f, dtype=float)
    cols = f.reshape(t.size, -1)
    out = np.empty_like(cols)
    for j in range(cols.shape[1]):
        out[:, j] = InterpolatedUnivariateSpline(t, cols[:, j], k=3).antiderivative(1)(t)
    return out.reshape(f.shape)


def squad(R, t_in, t_out):
    """Spherical quadrangle interpolation of rotors R [N, 4] from t_in onto t_out (unequal steps; ends reflected)."""
    R = np.asarray(R, dtype=float)
    t_in = np.asarray(t_in, dtype=float)
    t_out = np.asarray(t_out, dtype=float)
    if R.size == 0 or t_out.size == 0:
        return np.empty((0, 4))
    n = R.shape[0]
    roll = lambda a, k: np.roll(a, k, axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        step = _q.log(_q.mul(_q.inverse(R), roll(R, -1)))
        prev = _q.log(_q.mul(_q.inverse(roll(R, 1)), R))
        nxt = _q.log(_q.mul(_q.inverse(roll(R, -1)), roll(R, -2)))
        h = roll(t_in, -1) - t_in
        A = _q.mul(R, _q.exp((-step + prev * (h / (t_in - roll(t_in, 1)))[:, None]) * 0.25))
        B = _q.mul(roll(R, -1), _q.exp((nxt * (h / (roll(t_in, -2) - roll(t_in, -1)))[:, None] - step) * -0.25))
    last_next = _q.mul(_q.mul(R[-1], _q.inverse(R[-2])), R[-1])
    A[0] = R[0]
    A[-1] = R[-1]
    B[-2] = R[-1]
    B[-1] = last_next
    R_ip1 = roll(R, -1).copy()
    R_ip1[-1] = last_next
    t_ip1 = roll(t_in, -1).copy()
    t_ip1[-1] = t_in[-1] + (t_in[-1] - t_in[-2])
    i_out = np.clip(t_in.searchsorted(t_out, side="right") - 1, 0, n - 1)
    tau = (t_out - t_in[i_out]) / (t_ip1 - t_in)[i_out]

    def slerp(q1, q2, tau):
        flip = np.sum((q1 - q2) ** 2, axis=-1) > 2.0
        q2 = np.where(flip[:, None], -q2, q2)
        return _q.mul(_q.exp(tau[:, None] * _q.log(_q.mul(q2, _q.inverse(q1)))), q1)

    return slerp(slerp(R[i_out], R_ip1[i_out], tau), slerp(A[i_out], B[i_out], tau), 2 * tau * (1 - tau))


def integrate_angular_velocity(t, omega, R0=None, tolerance=1e-12):
    """R(t_i) with dR/dt = (1/2) omega R, R(t_0) = R0; omega tabulated [N, 3] -> not-a-knot cubic spline; adaptive
    Dormand-Prince 8(5,3), absolute tolerance `tolerance`, output at the samples."""
    from scipy.integrate import solve_ivp

    t = np.asarray(t, dtype=float)
    spl = CubicSpline(t, np.asarray(omega, dtype=float))
    y0 = _q.one.copy() if R0 is None else np.asarray(R0, dtype=float)

    def RHS(tt, y):
        w = spl(tt)
        return 0.5 * np.array(
            [
                -w[0] * y[1] - w[1] * y[2] - w[2] * y[3],
                w[0] * y[0] + w[1] * y[3] - w[2] * y[2],
                -w[0] * y[3] + w[1] * y[0] + w[2] * y[1],
                w[0] * y[2] - w[1] * y[1] + w[2] * y[0],
            ]
        )

    sol = solve_ivp(RHS, [t[0], t[-1]], y0, method="DOP853", t_eval=t, atol=tolerance, rtol=100 * np.finfo(float).eps)
    return sol.y.T.copy()


def minimal_rotation(R, t, iterations=2):
    """R -> R exp(gamma z / 2), gamma chosen so that the rotation about the frame's own z axis vanishes; repeated."""
    R = np.asarray(R, dtype=float)
    t = np.asarray(t, dtype=float)
    for _ in range(iterations):
        Rdot = _spline_derivative(R, t)
        halfgammadot = _q.mul(_q.mul(Rdot, _q.z), _q.conj(R))[..., 0]
        halfgamma = _spline_antiderivative(halfgammadot, t)
        zer = np.zeros_like(halfgamma)
        R = _q.mul(R, _q.exp(np.stack([zer, zer, zer, halfgamma], axis=-1)))
    return R


def angular_velocity(R, t):
    """omega = 2 Rdot R^-1 (vector part), Rdot by cubic splines."""
    R = np.asarray(R, dtype=float)
    Rdot = _spline_derivative(R, np.asarray(t, dtype=float))
    return (2 * _q.mul(Rdot, _q.conj(R)))[..., 1:]
