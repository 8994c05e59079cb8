"""Stand-in for `numpy-quaternion` so that the UNMODIFIED reference (/root/reference/scri) can be imported here.

TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy-quaternion (pyproject.toml of the reference pins >=2024.0.2) is not
vendored in /root/reference and not installable in this image.  This module restates the part of its API that scri calls
(the call sites are listed in SURVEY.md section 8c) on top of oracle/quat.py.  Rotor arrays are numpy OBJECT arrays of
`quaternion` instances: `np.dtype(np.quaternion)` is then `object`, which is all scri checks
(scri/waveform_base.py:344).  Only used by tests/golden/make_reference_vectors.py and tests that run in this container;
never by scri_b200/.
"""
import math

import numpy as np

from oracle import quat as _q

__version__ = "2024.0.13+oracle.shim"


class quaternion:
    """One quaternion (w, x, y, z); Hamilton product."""

    __slots__ = ("w", "x", "y", "z")

    def __init__(self, *args):
        if len(args) == 4:
            self.w, self.x, self.y, self.z = (float(a) for a in args)
        elif len(args) == 3:
            self.w = 0.0
            self.x, self.y, self.z = (float(a) for a in args)
        elif len(args) == 1 and isinstance(args[0], quaternion):
            o = args[0]
            self.w, self.x, self.y, self.z = o.w, o.x, o.y, o.z
        elif len(args) == 1:
            self.w, self.x, self.y, self.z = float(args[0]), 0.0, 0.0, 0.0
        elif len(args) == 0:
            self.w = self.x = self.y = self.z = 0.0
        else:
            raise TypeError("quaternion takes 0, 1, 3 or 4 real components")

    # -- components ------------------------------------------------------------------------------------------------
    @property
    def components(self):
        return np.array([self.w, self.x, self.y, self.z])

    @property
    def vec(self):
        return np.array([self.x, self.y, self.z])

    @vec.setter
    def vec(self, v):
        self.x, self.y, self.z = (float(a) for a in v)

    @property
    def real(self):
        return self.w

    @property
    def a(self):
        return complex(self.w, self.z)

    @property
    def b(self):
        return complex(self.y, self.x)

    # -- algebra ---------------------------------------------------------------------------------------------------
    def __mul__(self, o):
        if isinstance(o, quaternion):
            a, b, c, d = self.w, self.x, self.y, self.z
            e, f, g, h = o.w, o.x, o.y, o.z
            return quaternion(
                a * e - b * f - c * g - d * h,
                a * f + b * e + c * h - d * g,
                a * g - b * h + c * e + d * f,
                a * h + b * g - c * f + d * e,
            )
        if isinstance(o, np.ndarray):
            return NotImplemented
        if isinstance(o, (int, float, np.integer, np.floating)):
            return quaternion(self.w * o, self.x * o, self.y * o, self.z * o)
        return NotImplemented

    def __rmul__(self, o):
        if isinstance(o, (int, float, np.integer, np.floating)):
            return quaternion(self.w * o, self.x * o, self.y * o, self.z * o)
        return NotImplemented

    def __truediv__(self, o):
        if isinstance(o, quaternion):
            return self * o.inverse()
        if isinstance(o, np.ndarray):
            return NotImplemented
        if isinstance(o, (int, float, np.integer, np.floating)):
            return quaternion(self.w / o, self.x / o, self.y / o, self.z / o)
        return NotImplemented

    def __rtruediv__(self, o):
        if isinstance(o, (int, float, np.integer, np.floating)):
            return self.inverse() * o
        return NotImplemented

    def __add__(self, o):
        if isinstance(o, quaternion):
            return quaternion(self.w + o.w, self.x + o.x, self.y + o.y, self.z + o.z)
        if isinstance(o, np.ndarray):
            return NotImplemented
        if isinstance(o, (int, float, np.integer, np.floating)):
            return quaternion(self.w + o, self.x, self.y, self.z)
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, quaternion):
            return quaternion(self.w - o.w, self.x - o.x, self.y - o.y, self.z - o.z)
        if isinstance(o, np.ndarray):
            return NotImplemented
        if isinstance(o, (int, float, np.integer, np.floating)):
            return quaternion(self.w - o, self.x, self.y, self.z)
        return NotImplemented

    def __rsub__(self, o):
        return (-self) + o

    def __neg__(self):
        return quaternion(-self.w, -self.x, -self.y, -self.z)

    def __pos__(self):
        return quaternion(self)

    def __invert__(self):
        return self.inverse()

    def __abs__(self):
        return self.abs()

    def __pow__(self, p):
        if isinstance(p, (int, float, np.integer, np.floating)):
            return (self.log() * p).exp()
        return NotImplemented

    def __eq__(self, o):
        return isinstance(o, quaternion) and (self.w, self.x, self.y, self.z) == (o.w, o.x, o.y, o.z)

    def __ne__(self, o):
        return not self.__eq__(o)

    __hash__ = None

    def __repr__(self):
        return f"quaternion({self.w!r}, {self.x!r}, {self.y!r}, {self.z!r})"

    def norm(self):
        return self.w * self.w + self.x * self.x + self.y * self.y + self.z * self.z

    def abs(self):
        return math.sqrt(self.norm())

    absolute = abs

    def conjugate(self):
        return quaternion(self.w, -self.x, -self.y, -self.z)

    conj = conjugate

    def inverse(self):
        with np.errstate(invalid="ignore", divide="ignore"):
            n = np.float64(self.norm())
            return quaternion(self.w / n, -self.x / n, -self.y / n, -self.z / n)

    def normalized(self):
        with np.errstate(invalid="ignore", divide="ignore"):
            n = np.float64(self.abs())          # 0/0 -> nan as in the C implementation, no exception
            return quaternion(self.w / n, self.x / n, self.y / n, self.z / n)

    def x_parity_conjugate(self):
        return quaternion(self.w, self.x, -self.y, -self.z)

    def y_parity_conjugate(self):
        return quaternion(self.w, -self.x, self.y, -self.z)

    def z_parity_conjugate(self):
        return quaternion(self.w, -self.x, -self.y, self.z)

    def parity_conjugate(self):
        return quaternion(self)

    def x_parity_symmetric_part(self):
        return quaternion(self.w, self.x, 0.0, 0.0)

    def x_parity_antisymmetric_part(self):
        return quaternion(0.0, 0.0, self.y, self.z)

    def y_parity_symmetric_part(self):
        return quaternion(self.w, 0.0, self.y, 0.0)

    def y_parity_antisymmetric_part(self):
        return quaternion(0.0, self.x, 0.0, self.z)

    def z_parity_symmetric_part(self):
        return quaternion(self.w, 0.0, 0.0, self.z)

    def z_parity_antisymmetric_part(self):
        return quaternion(0.0, self.x, self.y, 0.0)

    def parity_symmetric_part(self):
        return quaternion(self)

    def parity_antisymmetric_part(self):
        return quaternion(0.0, 0.0, 0.0, 0.0)

    def exp(self):
        return quaternion(*_q.exp(self.components))

    def log(self):
        return quaternion(*_q.log(self.components))

    def sqrt(self):
        """Principal square root: for a unit rotor (1 + q) / |1 + q|; in general sqrt|q| times that of q/|q|."""
        n = self.abs()
        if n == 0.0:
            return quaternion(0.0, 0.0, 0.0, 0.0)
        u = self / n
        if abs(u.w + 1.0) < 1e-14 and abs(u.x) + abs(u.y) + abs(u.z) < 1e-14:
            return quaternion(0.0, math.sqrt(n), 0.0, 0.0)
        p = quaternion(u.w + 1.0, u.x, u.y, u.z)
        return p * (math.sqrt(n) / p.abs())

    def isnan(self):
        return any(math.isnan(c) for c in (self.w, self.x, self.y, self.z))


np.quaternion = quaternion

# numpy-quaternion registers ufunc loops for its dtype; object arrays have none for the predicates scri calls on rotor
# arrays (scri/waveform_base.py ensure_validity), so those three names get a front end that handles rotor arrays and
# hands everything else to the real ufunc.  Only processes that import this shim (golden generation, shim tests) see it.
class _rotor_predicate:
    def __init__(self, ufunc, combine):
        self.__wrapped__ = ufunc
        self._combine = combine
        self.__name__ = ufunc.__name__

    def __getattr__(self, name):          # nin, nout, reduce, at, ...: the real ufunc's
        return getattr(self.__wrapped__, name)

    def __call__(self, a, *args, **kwargs):
        ufunc, combine = self.__wrapped__, self._combine
        if isinstance(a, quaternion):
            return combine(ufunc(a.components))
        if isinstance(a, np.ndarray) and a.dtype == object:
            out = np.empty(a.shape, dtype=bool)
            flat = out.reshape(-1)
            for i, q in enumerate(a.reshape(-1)):
                flat[i] = combine(ufunc(q.components)) if isinstance(q, quaternion) else ufunc(q)
            return out
        return ufunc(a, *args, **kwargs)


def _elementwise(name):
    def f(a):
        if isinstance(a, quaternion):
            return getattr(a, name)()
        a = np.asarray(a, dtype=object)
        out = np.empty(a.shape, dtype=object)
        flat = out.reshape(-1)
        for i, q in enumerate(a.reshape(-1)):
            flat[i] = getattr(q, name)()
        return out

    f.__name__ = name
    return f


for _name in [p + k for p in ("x_parity_", "y_parity_", "z_parity_", "parity_") for k in ("conjugate", "symmetric_part", "antisymmetric_part")]:
    setattr(np, _name, _elementwise(_name))

if not hasattr(np.isfinite, "__wrapped__"):
    np.isfinite = _rotor_predicate(np.isfinite, np.all)
    np.isnan = _rotor_predicate(np.isnan, np.any)
    np.isinf = _rotor_predicate(np.isinf, np.any)

one = quaternion(1.0, 0.0, 0.0, 0.0)
x = quaternion(0.0, 1.0, 0.0, 0.0)
y = quaternion(0.0, 0.0, 1.0, 0.0)
z = quaternion(0.0, 0.0, 0.0, 1.0)
zero = quaternion(0.0, 0.0, 0.0, 0.0)


def as_float_array(a):
    """[..., 4] float array of the components."""
    if isinstance(a, quaternion):
        return a.components
    a = np.asarray(a, dtype=object)
    out = np.empty(a.shape + (4,))
    flat = out.reshape(-1, 4)
    for i, q in enumerate(a.reshape(-1)):
        flat[i] = (q.w, q.x, q.y, q.z)
    return out


def as_quat_array(a):
    """Object array of quaternions from a float array [..., 4]."""
    a = np.asarray(a, dtype=float)
    if a.shape[-1] != 4:
        raise ValueError(f"last dimension must be 4, not {a.shape}")
    out = np.empty(a.shape[:-1], dtype=object)
    flat = out.reshape(-1) if out.ndim else None
    if flat is None:
        return quaternion(*a)
    for i, c in enumerate(a.reshape(-1, 4)):
        flat[i] = quaternion(*c)
    return out


from_float_array = as_quat_array


def as_vector_part(a):
    return as_float_array(a)[..., 1:]


def from_vector_part(v, vector_axis=-1):
    v = np.asarray(v, dtype=float)
    return as_quat_array(np.concatenate([np.zeros(v.shape[:-1] + (1,)), v], axis=-1))


def as_spinor_array(a):
    """[..., 2] complex (w + i z, y + i x) -- scri/rotations.py:311."""
    return _q.as_spinor_array(as_float_array(a))


def from_spherical_coords(theta_phi, phi=None):
    if phi is None:
        theta_phi = np.asarray(theta_phi, dtype=float)
        theta, phi = theta_phi[..., 0], theta_phi[..., 1]
    else:
        theta = theta_phi
    r = _q.from_spherical_coords(theta, phi)
    return quaternion(*r) if r.ndim == 1 else as_quat_array(r)


def as_spherical_coords(q):
    return _q.as_spherical_coords(as_float_array(q))


def from_rotation_vector(rot):
    rot = np.asarray(rot, dtype=float)
    qa = np.zeros(rot.shape[:-1] + (4,))
    qa[..., 1:] = rot / 2
    r = _q.exp(qa)
    return quaternion(*r) if r.ndim == 1 else as_quat_array(r)


def as_rotation_vector(q):
    return 2 * _q.log(as_float_array(q))[..., 1:]


def from_euler_angles(alpha_beta_gamma, beta=None, gamma=None):
    if gamma is None:
        abg = np.asarray(alpha_beta_gamma, dtype=float)
        alpha, beta, gamma = abg[..., 0], abg[..., 1], abg[..., 2]
    else:
        alpha = alpha_beta_gamma
    r = _q.from_euler_angles(alpha, beta, gamma)
    return quaternion(*r) if r.ndim == 1 else as_quat_array(r)


def rotate_vectors(R, v, axis=-1):
    """Rotate the 3-vectors along `axis` of v by the rotor(s) R; result has shape R.shape + v.shape."""
    Rf = as_float_array(R)
    v = np.asarray(v, dtype=float)
    v = np.moveaxis(v, axis, -1)
    Rf2 = Rf.reshape(Rf.shape[:-1] + (1,) * (v.ndim - 1) + (4,))
    out = _q.rotate_vector(Rf2 / _q.absq(Rf2)[..., None], v)
    return np.moveaxis(out, -1, axis if axis >= 0 else axis)


def rotor_chordal_distance(p, q):
    return abs(p - q)


def rotor_intrinsic_distance(p, q):
    return 2 * abs((p.inverse() * q).log())


def slerp_evaluate(q1, q2, tau):
    """numpy-quaternion's slerp: (q2 / q1)^tau q1, through the shorter arc."""
    if (q1 - q2).norm() > 2.0:
        q2 = -q2
    return ((q2 / q1) ** tau) * q1


def squad_evaluate(tau, q_i, a_i, b_ip1, q_ip1):
    return slerp_evaluate(slerp_evaluate(q_i, q_ip1, tau), slerp_evaluate(a_i, b_ip1, tau), 2 * tau * (1 - tau))


from .calculus import (  # noqa: E402
    antiderivative, definite_integral, derivative, indefinite_integral, spline_definite_integral, spline_derivative,
    spline_indefinite_integral,
)
from .quaternion_time_series import (  # noqa: E402
    angular_velocity, integrate_angular_velocity, minimal_rotation, squad, unflip_rotors,
)
from . import calculus, means, quaternion_time_series  # noqa: E402,F401


def optimal_alignment_in_Euclidean_metric(a, b, t=None):
    """Rotor R minimising the integral of |R a R^-1 - b|^2 (Horn/Kabsch via the 4x4 symmetric matrix)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    if t is None:
        S = np.einsum("ij,ik->jk", a, b)
    else:
        S = spline_definite_integral(a[:, :, None] * b[:, None, :], t)
    N = np.array(
        [
            [S[0, 0] + S[1, 1] + S[2, 2], S[1, 2] - S[2, 1], S[2, 0] - S[0, 2], S[0, 1] - S[1, 0]],
            [S[1, 2] - S[2, 1], S[0, 0] - S[1, 1] - S[2, 2], S[0, 1] + S[1, 0], S[2, 0] + S[0, 2]],
            [S[2, 0] - S[0, 2], S[0, 1] + S[1, 0], -S[0, 0] + S[1, 1] - S[2, 2], S[1, 2] + S[2, 1]],
            [S[0, 1] - S[1, 0], S[2, 0] + S[0, 2], S[1, 2] + S[2, 1], -S[0, 0] - S[1, 1] + S[2, 2]],
        ]
    )
    w, v = np.linalg.eigh(N)
    return quaternion(*v[:, -1])
