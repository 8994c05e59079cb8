"""quaternion.quaternion_time_series as scri calls it (mode_calculations.py:464-467, rotations.py:38-43,
waveform_base.py:634-646,962).  TEST INFRASTRUCTURE: thin object-array front end of oracle/quat_series.py, where the
algorithms are restated and their pinning status is described."""
import numpy as np

from oracle import quat_series as _qs


def _Q():
    import quaternion

    return quaternion


def unflip_rotors(q, axis=-1, inplace=False):
    Q = _Q()
    f = Q.as_float_array(q).copy()
    d = np.sum(f[1:] * f[:-1], axis=-1)
    sign = np.cumprod(np.where(d < 0, -1.0, 1.0), axis=0)
    f[1:] *= sign[..., None]
    return Q.as_quat_array(f)


def squad(R_in, t_in, t_out, unflip_input_rotors=False):
    Q = _Q()
    R_in = np.asarray(R_in, dtype=object)
    if R_in.size == 0 or np.size(t_out) == 0:
        return np.array((), dtype=object)
    if unflip_input_rotors:
        R_in = unflip_rotors(R_in, axis=0)
    return Q.as_quat_array(_qs.squad(Q.as_float_array(R_in), t_in, t_out))


def integrate_angular_velocity(Omega, t0, t1, R0=None, tolerance=1e-12):
    Q = _Q()
    t_Omega, v = Omega
    y0 = None if R0 is None else (R0.components if isinstance(R0, Q.quaternion) else np.asarray(R0, dtype=float))
    R = _qs.integrate_angular_velocity(t_Omega, v, y0, tolerance)
    return np.asarray(t_Omega, dtype=float), Q.as_quat_array(R)


def minimal_rotation(R, t, iterations=2):
    Q = _Q()
    return Q.as_quat_array(_qs.minimal_rotation(Q.as_float_array(R), t, iterations))


def angular_velocity(R, t):
    return _qs.angular_velocity(_Q().as_float_array(R), t)
