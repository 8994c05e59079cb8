"""spherical_functions.Grid as scri/asymptotic_bondi_data/transformations.py uses it: function values on the
(theta, phi) grid (last two axes) carrying a spin weight; products add the weights, real/imag/conj keep track of it.
TEST INFRASTRUCTURE (see oracle/refshim/spherical_functions/__init__.py)."""
import copy

import numpy as np


class Grid(np.ndarray):
    def __new__(cls, input_array, *args, **kwargs):
        metadata = copy.copy(getattr(input_array, "_metadata", {}))
        metadata.update(**kwargs)
        arr = np.asanyarray(input_array)          # the dtype is kept: real grids (conformal factor, alpha) stay real
        if metadata.get("spin_weight", None) is None:
            raise ValueError("Spin weight must be specified")
        obj = arr.view(cls)
        obj._metadata = dict(metadata)
        obj._metadata["spin_weight"] = int(metadata["spin_weight"])
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self._metadata = copy.copy(getattr(obj, "_metadata", {}))

    @property
    def s(self):
        return self._metadata["spin_weight"]

    spin_weight = s

    @property
    def ndarray(self):
        return np.ndarray.view(self, np.ndarray)

    @property
    def n_theta(self):
        return self.shape[-2]

    @property
    def n_phi(self):
        return self.shape[-1]

    def _wrap(self, a, s):
        out = np.asarray(a).view(Grid)
        out._metadata = {"spin_weight": int(s)}
        return out

    def __getitem__(self, key):
        res = np.ndarray.__getitem__(self.ndarray, key)
        if isinstance(res, np.ndarray) and res.ndim >= 2:
            return self._wrap(res, self.s)
        return res

    def __setitem__(self, key, value):
        np.ndarray.__setitem__(self.ndarray, key, np.asarray(value))

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        ss = [i.s if isinstance(i, Grid) else None for i in inputs]
        ins = [i.ndarray if isinstance(i, Grid) else np.asarray(i) for i in inputs]
        # an operand that is not a grid broadcasts against the leading (e.g. time) axes, as in sf
        shape2 = next(i.shape[-2:] for i in inputs if isinstance(i, Grid))
        ins = [
            a[..., np.newaxis, np.newaxis] if (s_ is None and a.ndim > 0 and a.shape[-2:] != shape2) else a
            for a, s_ in zip(ins, ss)
        ]
        if out is not None:
            kwargs["out"] = tuple(o.ndarray if isinstance(o, Grid) else o for o in out)
        res = getattr(ufunc, method)(*ins, **kwargs)
        if method != "__call__" or not isinstance(res, np.ndarray) or res.ndim < 2:
            return res
        name = ufunc.__name__
        if name == "multiply":
            s = sum(x for x in ss if x is not None)
        elif name in ("divide", "true_divide"):
            s = (ss[0] or 0) - (ss[1] or 0)
        elif name in ("add", "subtract"):
            known = [x for x in ss if x is not None]
            if len(set(known)) > 1:
                raise ValueError(f"Cannot {name} grids of different spin weights {known}")
            s = known[0]
        elif name in ("conjugate", "conj"):
            s = -ss[0]
        elif name in ("negative", "positive"):
            s = ss[0]
        elif name == "power":
            s = ss[0] * int(np.asarray(ins[1]).flat[0]) if ss[0] else 0
        elif name in ("absolute", "sqrt", "exp", "log", "reciprocal"):
            if name == "reciprocal":
                s = -ss[0]
            else:
                return res
        else:
            return res
        return self._wrap(res, s)

    @property
    def real(self):
        return self._wrap(self.ndarray.real, self.s)

    @property
    def imag(self):
        return self._wrap(self.ndarray.imag, self.s)

    @property
    def bar(self):
        return self._wrap(np.conj(self.ndarray), -self.s)

    def modes(self, ell_max=None, **kwargs):
        import spinsfast

        from .SWSH_modes import Modes

        if ell_max is None:
            ell_max = (min(self.n_theta, self.n_phi) - 1) // 2
        return Modes(spinsfast.map2salm(self.ndarray, self.s, ell_max), spin_weight=self.s, ell_max=ell_max)
