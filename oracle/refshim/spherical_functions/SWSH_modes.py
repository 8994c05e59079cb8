"""spherical_functions.Modes as scri's ModesTimeSeries / AsymptoticBondiData use it.  TEST INFRASTRUCTURE (see
oracle/refshim/spherical_functions/__init__.py): an ndarray subclass whose last axis holds the mode weights from ell = 0,
with the spin weight, ell_max and the multiplication truncator in `_metadata`.  Algebra follows the sf documentation:
bar, real, imag, eth / ethbar (Newman-Penrose normalisation), multiply (Wigner-3j sum), evaluate, grid."""
import copy
import math

import numpy as np

from oracle import sf as _sf


def _ells(ell_max):
    return np.array([ell for ell in range(ell_max + 1) for _ in range(2 * ell + 1)], dtype=float)


class Modes(np.ndarray):
    def __new__(cls, input_array, *args, **kwargs):
        if len(args) > 0:
            raise ValueError("Modes takes keyword arguments only besides the array")
        metadata = copy.copy(getattr(input_array, "_metadata", {}))
        metadata.update(**kwargs)
        arr = np.asanyarray(input_array)
        if arr.dtype != complex:
            arr = arr.astype(complex)
        s = metadata.get("spin_weight", None)
        if s is None:
            raise ValueError("Spin weight must be specified")
        ell_min = metadata.get("ell_min", 0)
        ell_max = metadata.get("ell_max", None)
        n = arr.shape[-1]
        if ell_max is None:
            ell_max = int(round(math.sqrt(n + ell_min**2))) - 1
        if (ell_max + 1) ** 2 - ell_min**2 != n:
            raise ValueError(f"last axis has size {n}, not (ell_max+1)^2 - ell_min^2 for ell_min={ell_min}, ell_max={ell_max}")
        if ell_min != 0:
            arr = np.concatenate([np.zeros(arr.shape[:-1] + (ell_min**2,), dtype=complex), arr], axis=-1)
        obj = arr.view(cls)
        obj._metadata = dict(metadata)
        obj._metadata["spin_weight"] = int(s)
        obj._metadata["ell_min"] = 0
        obj._metadata["ell_max"] = int(ell_max)
        obj._metadata.setdefault("multiplication_truncator", max)
        if abs(s) > 0:
            np.ndarray.view(obj, np.ndarray)[..., : int(s) ** 2] = 0.0
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self._metadata = copy.copy(getattr(obj, "_metadata", {}))

    # numpy ufuncs and reductions see plain arrays; the algebra below re-wraps explicitly
    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        ins = [np.ndarray.view(i, np.ndarray) if isinstance(i, Modes) else i for i in inputs]
        if out is not None:
            kwargs["out"] = tuple(np.ndarray.view(o, np.ndarray) if isinstance(o, Modes) else o for o in out)
        res = getattr(ufunc, method)(*ins, **kwargs)
        if out is not None and len(out) == 1 and isinstance(out[0], Modes):
            return out[0]
        return res

    # -- metadata ----------------------------------------------------------------------------------------------
    @property
    def s(self):
        return self._metadata["spin_weight"]

    spin_weight = s

    @property
    def ell_min(self):
        return 0

    @property
    def ell_max(self):
        return self._metadata["ell_max"]

    @property
    def multiplication_truncator(self):
        return self._metadata["multiplication_truncator"]

    @property
    def ndarray(self):
        return np.ndarray.view(self, np.ndarray)

    def index(self, ell, m):
        return ell * (ell + 1) + m

    def _wrap(self, array, **updates):
        out = np.ndarray.view(np.ascontiguousarray(array), type(self))
        out._metadata = copy.copy(self._metadata)
        out._metadata.update(updates)
        return out

    def __getitem__(self, key):
        res = np.ndarray.__getitem__(self.ndarray, key)
        if isinstance(res, np.ndarray) and res.ndim >= 1 and res.shape[-1] == self.shape[-1] and res.ndim == self.ndim:
            return self._wrap(res) if not res.flags.c_contiguous else self._view_like(res)
        return res

    def _view_like(self, res):
        out = res.view(type(self))
        out._metadata = copy.copy(self._metadata)
        return out

    def __setitem__(self, key, value):
        np.ndarray.__setitem__(self.ndarray, key, np.asarray(value))

    def truncate_ell(self, new_ell_max):
        if new_ell_max >= self.ell_max:
            return self._wrap(self.ndarray.copy())
        return self._wrap(self.ndarray[..., : (new_ell_max + 1) ** 2].copy(), ell_max=new_ell_max)

    def _padded(self, ell_max):
        a = self.ndarray
        if ell_max == self.ell_max:
            return a
        if ell_max < self.ell_max:
            return a[..., : (ell_max + 1) ** 2]
        return np.concatenate([a, np.zeros(a.shape[:-1] + ((ell_max + 1) ** 2 - a.shape[-1],), dtype=complex)], axis=-1)

    # -- the function's algebra ----------------------------------------------------------------------------------
    @property
    def bar(self):
        """Conjugate of the function: fbar_{l,m} = (-1)^{s+m} conj(f_{l,-m}); spin weight -s."""
        a = self.ndarray
        out = np.empty_like(a)
        s = self.s
        for ell in range(self.ell_max + 1):
            lo = ell * ell
            m = np.arange(-ell, ell + 1)
            sign = np.where((s + m) % 2 == 0, 1.0, -1.0)
            out[..., lo : lo + 2 * ell + 1] = sign * np.conj(a[..., lo : lo + 2 * ell + 1][..., ::-1])
        return self._wrap(out, spin_weight=-s)

    conjugate = conj = property(lambda self: self.bar)

    @property
    def real(self):
        if self.s != 0:
            raise ValueError("The real part of a function with nonzero spin weight is not a spin-weighted function")
        return self._wrap(0.5 * (self.ndarray + self.bar.ndarray))

    @property
    def imag(self):
        if self.s != 0:
            raise ValueError("The imaginary part of a function with nonzero spin weight is not a spin-weighted function")
        return self._wrap((self.ndarray - self.bar.ndarray) / 2j)

    @property
    def eth(self):
        ell = _ells(self.ell_max)
        s = self.s
        fac = np.where(ell >= abs(s + 1), np.sqrt(np.maximum((ell - s) * (ell + s + 1), 0.0)), 0.0)
        return self._wrap(self.ndarray * fac, spin_weight=s + 1)

    @property
    def ethbar(self):
        ell = _ells(self.ell_max)
        s = self.s
        fac = np.where(ell >= abs(s - 1), -np.sqrt(np.maximum((ell + s) * (ell - s + 1), 0.0)), 0.0)
        return self._wrap(self.ndarray * fac, spin_weight=s - 1)

    @property
    def eth_GHP(self):
        return self.eth / math.sqrt(2)

    @property
    def ethbar_GHP(self):
        return self.ethbar / math.sqrt(2)

    def norm(self):
        return np.linalg.norm(self.ndarray, axis=-1)

    def evaluate(self, rotors):
        import quaternion
        import spherical_functions as sf

        Rf = quaternion.as_float_array(rotors)
        Y = sf.SWSH_grid(quaternion.as_quat_array(Rf.reshape(-1, 4)), self.s, self.ell_max)  # [points, lm]
        vals = np.tensordot(self.ndarray, Y, axes=([-1], [-1]))
        return vals.reshape(self.shape[:-1] + Rf.shape[:-1])

    def grid(self, n_theta=None, n_phi=None, **kwargs):
        import spinsfast

        from .SWSH_grids import Grid

        n_theta = n_theta or 2 * self.ell_max + 1
        n_phi = n_phi or n_theta
        return Grid(spinsfast.salm2map(self.ndarray, self.s, self.ell_max, n_theta, n_phi), spin_weight=self.s)

    def multiply(self, other, truncator=None):
        if not isinstance(other, Modes):
            return self * other
        truncator = truncator or self.multiplication_truncator
        ell_out = int(truncator((self.ell_max, other.ell_max)))
        a, b = self.ndarray, other.ndarray
        lead = np.broadcast_shapes(a.shape[:-1], b.shape[:-1])
        a = np.broadcast_to(a, lead + a.shape[-1:])
        b = np.broadcast_to(b, lead + b.shape[-1:])
        out = _sf.modes_multiply(a, self.s, self.ell_max, b, other.s, other.ell_max, ell_out)
        return self._wrap(out, spin_weight=self.s + other.s, ell_max=ell_out)

    # -- arithmetic ----------------------------------------------------------------------------------------------
    def _binary_sum(self, other, sign):
        if isinstance(other, Modes):
            if other.s != self.s:
                raise ValueError(f"Cannot add modes with different spin weights ({self.s} and {other.s})")
            L = max(self.ell_max, other.ell_max)
            return self._wrap(self._padded(L) + sign * other._padded(L), ell_max=L)
        # a plain array is taken as mode weights of the same spin weight (sf broadcasts it against the data)
        return self._wrap(self.ndarray + sign * np.asarray(other))

    def __add__(self, other):
        return self._binary_sum(other, 1.0)

    __radd__ = __add__

    def __sub__(self, other):
        return self._binary_sum(other, -1.0)

    def __rsub__(self, other):
        return (-self)._binary_sum(other, 1.0)

    def __neg__(self):
        return self._wrap(-self.ndarray)

    def __pos__(self):
        return self._wrap(self.ndarray.copy())

    def __mul__(self, other):
        if isinstance(other, Modes):
            return self.multiply(other)
        other = np.asarray(other)
        if other.ndim > 0:            # arrays multiply function values: they broadcast against all but the mode axis
            other = other[..., np.newaxis]
        return self._wrap(self.ndarray * other)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Modes):
            raise ValueError("Cannot divide one Modes object by another")
        other = np.asarray(other)
        if other.ndim > 0:
            other = other[..., np.newaxis]
        return self._wrap(self.ndarray / other)

    def __repr__(self):
        return f"{type(self).__name__}({self.ndarray!r}, spin_weight={self.s}, ell_max={self.ell_max})"
