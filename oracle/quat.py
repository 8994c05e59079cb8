"""Restatement of the numpy-quaternion operations the scri hot path uses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Third-party package `numpy-quaternion`
(pyproject.toml:22 of the reference pins >=2024.0.2) is not vendored in /root/reference;
the call sites followed here are scri/waveform_grid.py:113-174,472, scri/rotations.py:311,
scri/mode_calculations.py:429-486, scri/sample_waveforms.py:383-533.

Quaternions are float arrays [..., 4] = (w, x, y, z), Hamilton product.
"""
import numpy as np

one = np.array([1.0, 0.0, 0.0, 0.0])
x = np.array([0.0, 1.0, 0.0, 0.0])
y = np.array([0.0, 0.0, 1.0, 0.0])
z = np.array([0.0, 0.0, 0.0, 1.0])


def mul(p, q):
    p = np.asarray(p, dtype=float)
    q = np.asarray(q, dtype=float)
    pw, px, py, pz = p[..., 0], p[..., 1], p[..., 2], p[..., 3]
    qw, qx, qy, qz = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack(
        [
            pw * qw - px * qx - py * qy - pz * qz,
            pw * qx + px * qw + py * qz - pz * qy,
            pw * qy - px * qz + py * qw + pz * qx,
            pw * qz + px * qy - py * qx + pz * qw,
        ],
        axis=-1,
    )


def conj(q):
    q = np.asarray(q, dtype=float)
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def norm2(q):
    q = np.asarray(q, dtype=float)
    return np.sum(q * q, axis=-1)


def absq(q):
    return np.sqrt(norm2(q))


def normalized(q):
    q = np.asarray(q, dtype=float)
    return q / absq(q)[..., None]


def inverse(q):
    return conj(q) / norm2(q)[..., None]


def exp(q):
    """Quaternion exponential."""
    q = np.asarray(q, dtype=float)
    v = q[..., 1:]
    vn = np.sqrt(np.sum(v * v, axis=-1))
    ew = np.exp(q[..., 0])
    with np.errstate(invalid="ignore", divide="ignore"):
        s = np.where(vn > 0, np.sin(vn) / np.where(vn > 0, vn, 1.0), 1.0)
    out = np.empty(q.shape)
    out[..., 0] = ew * np.cos(vn)
    out[..., 1:] = (ew * s)[..., None] * v
    return out


def log(q):
    q = np.asarray(q, dtype=float)
    v = q[..., 1:]
    vn = np.sqrt(np.sum(v * v, axis=-1))
    n = absq(q)
    out = np.zeros(q.shape)
    out[..., 0] = np.log(n)
    ang = np.arctan2(vn, q[..., 0])
    with np.errstate(invalid="ignore", divide="ignore"):
        f = np.where(vn > 0, ang / np.where(vn > 0, vn, 1.0), 0.0)
    out[..., 1:] = f[..., None] * v
    return out


def sqrt(q):
    """Square root of a unit rotor: (1+q)/|1+q| (q != -1)."""
    q = np.asarray(q, dtype=float)
    p = q + one
    return p / absq(p)[..., None] * np.sqrt(absq(q))[..., None]


def from_spherical_coords(theta, phi):
    """exp(phi z/2) exp(theta y/2)"""
    theta = np.asarray(theta, dtype=float)
    phi = np.asarray(phi, dtype=float)
    ct, st = np.cos(theta / 2), np.sin(theta / 2)
    cp, sp = np.cos(phi / 2), np.sin(phi / 2)
    return np.stack([cp * ct, -sp * st, cp * st, sp * ct], axis=-1)


def as_spherical_coords(q):
    """(theta, phi) of the rotor, as in quaternion.as_spherical_coords (euler beta, alpha)."""
    q = np.asarray(q, dtype=float)
    n = norm2(q)
    theta = 2 * np.arccos(np.sqrt((q[..., 0] ** 2 + q[..., 3] ** 2) / n))
    phi = np.arctan2(q[..., 3], q[..., 0]) + np.arctan2(-q[..., 1], q[..., 2])
    return np.stack([theta, phi], axis=-1)


def from_euler_angles(alpha, beta, gamma):
    alpha, beta, gamma = (np.asarray(a, dtype=float) for a in (alpha, beta, gamma))
    return np.stack(
        [
            np.cos(beta / 2) * np.cos((alpha + gamma) / 2),
            -np.sin(beta / 2) * np.sin((alpha - gamma) / 2),
            np.sin(beta / 2) * np.cos((alpha - gamma) / 2),
            np.cos(beta / 2) * np.sin((alpha + gamma) / 2),
        ],
        axis=-1,
    )


def as_spinor_array(q):
    """[..., 2] complex (Ra, Rb) = (w + i z, y + i x)."""
    q = np.asarray(q, dtype=float)
    return np.stack([q[..., 0] + 1j * q[..., 3], q[..., 2] + 1j * q[..., 1]], axis=-1)


def rotate_vector(R, v):
    """R v R^-1 for unit (or not) rotor R and 3-vector(s) v."""
    v = np.asarray(v, dtype=float)
    vq = np.concatenate([np.zeros(v.shape[:-1] + (1,)), v], axis=-1)
    return mul(mul(R, vq), inverse(R))[..., 1:]


def rotor_intrinsic_distance(p, q):
    return 2 * absq(log(mul(inverse(p), q)))
