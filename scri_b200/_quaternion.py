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
