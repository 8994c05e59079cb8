"""Minimal rotor algebra on float arrays [..., 4] = (w, x, y, z) for the host side of the B200 path.

The reference passes `numpy-quaternion` objects (scri/waveform_grid.py:113-174, rotations.py:284-343);
that package is optional here: `as_float_quat` accepts np.quaternion arrays when it is importable and
plain float arrays otherwise.
"""
import numpy as np


def as_float_quat(q):
    """Float array [..., 4] from a float array / list / np.quaternion (array)."""
    if type(q).__name__ == "quaternion" or (isinstance(q, np.ndarray) and q.dtype.name == "quaternion"):
        import quaternion  # noqa: F401

        return quaternion.as_float_array(q)
    a = np.asarray(q)
    if a.dtype == object:
        import quaternion

        return quaternion.as_float_array(np.asarray(q, dtype=np.quaternion))
    a = np.asarray(a, dtype=float)
    if a.shape[-1:] != (4,):
        raise ValueError(f"Quaternion arrays must have last dimension 4; got shape {a.shape}")
    return a


def qmul(p, q):
    p = np.asarray(p, dtype=float)
    q = np.asarray(q, dtype=float)
    a, b, c, d = (p[..., i] for i in range(4))
    e, f, g, h = (q[..., i] for i in range(4))
    return np.stack(
        [
            a * e - b * f - c * g - d * h,
            a * f + b * e + c * h - d * g,
            a * g - b * h + c * e + d * f,
            a * h + b * g - c * f + d * e,
        ],
        axis=-1,
    )


def qconj(q):
    q = np.asarray(q, dtype=float)
    out = q.copy()
    out[..., 1:] *= -1
    return out


def qabs(q):
    return np.sqrt(np.sum(np.asarray(q, dtype=float) ** 2, axis=-1))


def qnormalized(q):
    q = np.asarray(q, dtype=float)
    return q / qabs(q)[..., None]


def qinverse(q):
    q = np.asarray(q, dtype=float)
    return qconj(q) / np.sum(q * q, axis=-1)[..., None]


def qexp_vec(v):
    """exp of the pure-vector quaternion (0, v)."""
    v = np.asarray(v, dtype=float)
    n = np.sqrt(np.sum(v * v, axis=-1))
    out = np.empty(v.shape[:-1] + (4,))
    out[..., 0] = np.cos(n)
    s = np.where(n > 0, np.sin(n) / np.where(n > 0, n, 1.0), 1.0)
    out[..., 1:] = s[..., None] * v
    return out


def qlog(q):
    q = np.asarray(q, dtype=float)
    v = q[..., 1:]
    vn = np.sqrt(np.sum(v * v, axis=-1))
    out = np.zeros(q.shape)
    out[..., 0] = np.log(qabs(q))
    f = np.where(vn > 0, np.arctan2(vn, q[..., 0]) / np.where(vn > 0, vn, 1.0), 0.0)
    out[..., 1:] = f[..., None] * v
    return out


def qsqrt(q):
    """Principal square root of unit rotors, (1+q)/|1+q|."""
    q = np.asarray(q, dtype=float)
    p = q.copy()
    p[..., 0] += 1.0
    return p / qabs(p)[..., None]


def from_spherical_coords(theta, phi):
    theta = np.asarray(theta, dtype=float)
    phi = np.asarray(phi, dtype=float)
    ct, st = np.cos(theta / 2), np.sin(theta / 2)
    cp, sp = np.cos(phi / 2), np.sin(phi / 2)
    return np.stack([cp * ct, -sp * st, cp * st, sp * ct], axis=-1)


def as_spherical_coords(q):
    q = np.asarray(q, dtype=float)
    n = np.sum(q * q, axis=-1)
    theta = 2 * np.arccos(np.sqrt((q[..., 0] ** 2 + q[..., 3] ** 2) / n))
    phi = np.arctan2(q[..., 3], q[..., 0]) + np.arctan2(-q[..., 1], q[..., 2])
    return theta, phi


def as_spinor_array(q):
    """[..., 2] complex: (Ra, Rb) = (w + i z, y + i x)   (quaternion.as_spinor_array; rotations.py:311)."""
    q = np.asarray(q, dtype=float)
    return np.stack([q[..., 0] + 1j * q[..., 3], q[..., 2] + 1j * q[..., 1]], axis=-1)


def rotate_z(R):
    """Vector part of R z R^-1 for rotors R[..., 4]."""
    R = np.asarray(R, dtype=float)
    w, x, y, z = (R[..., i] for i in range(4))
    n = w * w + x * x + y * y + z * z
    return np.stack([2 * (x * z + w * y), 2 * (y * z - w * x), w * w - x * x - y * y + z * z], axis=-1) / n[..., None]


def qexp(q):
    """exp of a general quaternion (w, v) = e^w (cos|v|, sin|v| v/|v|)."""
    q = np.asarray(q, dtype=float)
    return np.exp(q[..., :1]) * qexp_vec(q[..., 1:])


def slerp(q1, q2, tau):
    """(q2 q1^-1)^tau q1 - the geodesic from q1 (tau = 0) to q2 (tau = 1); `tau` broadcasts against the leading axes."""
    q1, q2 = np.asarray(q1, dtype=float), np.asarray(q2, dtype=float)
    tau = np.asarray(tau, dtype=float)[..., None]
    # numpy-quaternion's slerp goes the short way round: -q2 stands in for q2 when the rotors are more than sqrt(2) apart
    far = np.sum((q1 - q2) ** 2, axis=-1, keepdims=True) > 2.0
    q2 = np.where(far, -q2, q2)
    return qmul(qexp(tau * qlog(qmul(q2, qinverse(q1)))), q1)


def squad(R_in, t_in, t_out):
    """Spherical "quadrangle" interpolation of a rotor series onto new times: the C^1 analogue of a cubic spline on the
    rotation group (Shoemake 1987), with the control points corrected for unequal time steps as numpy-quaternion's
    `squad` does it - the routine scri/waveform_base.py:962 uses for the frame in `interpolate`.

      A_i     = R_i     exp( (log(R_{i-1}^-1 R_i) h_i / h_{i-1} - log(R_i^-1 R_{i+1})) / 4 )
      B_{i+1} = R_{i+1} exp(-(log(R_{i+1}^-1 R_{i+2}) h_i / h_{i+1} - log(R_i^-1 R_{i+1})) / 4 )
      R(t)    = slerp( slerp(R_i, R_{i+1}, tau), slerp(A_i, B_{i+1}, tau), 2 tau (1 - tau) ),  tau = (t - t_i) / h_i,

    the series continued at both ends by reflection (R_{-1} = R_0 R_1^-1 R_0, ...), which makes A_0 = R_0 and A_{n-1} = B_{n-1} =
    R_{n-1}.  Rotor arrays are float [..., 4] (w, x, y, z).  Checked against the output of the reference's own
    `interpolate` (tests/golden/reference_modes.npz, reference_frames.npz)."""
    R_in = as_float_quat(R_in)
    t_in = np.asarray(t_in, dtype=float)
    t_out = np.asarray(t_out, dtype=float)
    if R_in.size == 0 or t_out.size == 0:
        return np.empty((0, 4))
    n = R_in.shape[0]
    if n == 1:
        return np.repeat(R_in, t_out.size, axis=0)
    roll = lambda x, k: np.roll(x, k, axis=0)
    h = roll(t_in, -1) - t_in                                  # h_i = t_{i+1} - t_i (last entry wraps: overwritten below)
    with np.errstate(divide="ignore", invalid="ignore"):
        step = qlog(qmul(qinverse(R_in), roll(R_in, -1)))      # log(R_i^-1 R_{i+1})
        prev = qlog(qmul(qinverse(roll(R_in, 1)), R_in))       # log(R_{i-1}^-1 R_i)
        nxt = qlog(qmul(qinverse(roll(R_in, -1)), roll(R_in, -2)))   # log(R_{i+1}^-1 R_{i+2})
        A = qmul(R_in, qexp((prev * (h / (t_in - roll(t_in, 1)))[:, None] - step) * 0.25))
        B = qmul(roll(R_in, -1), qexp((nxt * (h / (roll(t_in, -2) - roll(t_in, -1)))[:, None] - step) * -0.25))
    last_next = qmul(qmul(R_in[-1], qinverse(R_in[-2])), R_in[-1])    # the reflected sample after the last one
    A[0] = R_in[0]
    A[-1] = R_in[-1]
    B[-2] = R_in[-1]
    B[-1] = last_next
    R_ip1 = roll(R_in, -1).copy()
    R_ip1[-1] = last_next
    t_ip1 = roll(t_in, -1).copy()
    t_ip1[-1] = t_in[-1] + (t_in[-1] - t_in[-2])
    i = np.clip(t_in.searchsorted(t_out, side="right") - 1, 0, n - 1)
    tau = (t_out - t_in[i]) / (t_ip1 - t_in)[i]
    return slerp(slerp(R_in[i], R_ip1[i], tau), slerp(A[i], B[i], tau), 2 * tau * (1 - tau))
