"""WaveformBase: the container shell shared by WaveformModes and WaveformGrid.

Host-side bookkeeping only (time axis, frame, provenance history, validity checks), mirroring the
public surface of scri/waveform_base.py: attribute names, `copy`, `_append_history` and the
`waveform_alterations` history-depth protocol (:38-62, :706-729), weights (:440-459), `norm` (:535-551),
`data_dot/ddot/int/iint` (:689-703) and `interpolate` (:949-967).  The numerical methods run on the GPU
through scri_b200.ops; there is no CPU fallback.
"""
import copy as _copy
import functools
import pprint
import warnings

import numpy as np

from . import _quaternion as Q
from .constants import (
    ConformalWeights,
    DataNames,
    FrameNames,
    MScaling,
    RScaling,
    SpinWeights,
    UnknownDataType,
    UnknownFrameType,
)


def waveform_alterations(func):
    """Keep the history nesting depth consistent around a mutating method (scri/waveform_base.py:38-62)."""

    @functools.wraps(func)
    def func_wrapper(self, *args, **kwargs):
        if self.__history_depth__ == 0:
            self._append_history("")
        stored = self.__history_depth__
        self.__history_depth__ += 1
        result = func(self, *args, **kwargs)
        self.__history_depth__ = stored
        return result

    return func_wrapper


class WaveformBase:
    __num = 0

    def __init__(self, *args, **kwargs):
        original_kwargs = kwargs.copy()
        override = kwargs.pop("override_exception_from_invalidity", False)
        self.__num = WaveformBase.__num
        WaveformBase.__num += 1
        self.__history_depth__ = 0
        if len(args) == 0:
            self.t = np.asarray(kwargs.pop("t", np.empty((0,), dtype=float)), dtype=float)
            frame = kwargs.pop("frame", np.empty((0, 4), dtype=float))
            self.frame = Q.as_float_quat(frame) if np.size(frame) else np.empty((0, 4), dtype=float)
            self.data = kwargs.pop("data", np.empty((0, 0), dtype=complex))
            self.history = list(kwargs.pop("history", []))
            self.version_hist = list(kwargs.pop("version_hist", []))
            self.frameType = kwargs.pop("frameType", UnknownFrameType)
            self.dataType = kwargs.pop("dataType", UnknownDataType)
            self.r_is_scaled_out = kwargs.pop("r_is_scaled_out", False)
            self.m_is_scaled_out = kwargs.pop("m_is_scaled_out", False)
            if "constructor_statement" in kwargs:
                self._append_history("{} = {}".format(self, kwargs.pop("constructor_statement")))
            else:
                opts = np.get_printoptions()
                np.set_printoptions(threshold=6)
                self._append_history("{} = {}(**{})".format(self, type(self).__name__, pprint.pformat(original_kwargs, indent=4)))
                np.set_printoptions(**opts)
        elif len(args) == 1 and isinstance(args[0], type(self)):
            other = args[0]
            self.t = np.copy(other.t)
            self.frame = np.copy(other.frame)
            self.data = np.copy(other.data)
            self.history = other.history[:]
            self.version_hist = other.version_hist[:]
            self.frameType = other.frameType
            self.dataType = other.dataType
            self.r_is_scaled_out = other.r_is_scaled_out
            self.m_is_scaled_out = other.m_is_scaled_out
            self._append_history(["", "{} = {}({})".format(self, type(self).__name__, other)])
        else:
            raise ValueError(
                f"Did not understand input arguments to `{type(self).__name__}` constructor.\n"
                "Note that explicit data values must be passed as keywords,\n"
                "whereas objects to be copied must be passed as the sole argument."
            )
        self._init_extra(args, kwargs)
        self.__history_depth__ = 1
        self.ensure_validity(alter=True, assertions=(not override))
        self.__history_depth__ = 0
        if kwargs:
            warnings.warn(
                f"\nIn `{type(self).__name__}` initializer, unused keyword arguments:\n" + pprint.pformat(kwargs, indent=4)
            )

    def _init_extra(self, args, kwargs):
        pass

    # ------------------------------------------------------------------ validity (waveform_base.py:272-424)
    def ensure_validity(self, alter=True, assertions=False):
        errors = []
        if not isinstance(self.t, np.ndarray) or self.t.dtype != np.dtype(float) or self.t.ndim != 1:
            errors.append("`t` must be a 1-d float numpy array")
        elif self.t.size > 1 and not (np.diff(self.t) > 0).all():
            errors.append("`t` must be strictly increasing")
        if not isinstance(self.data, np.ndarray):
            errors.append("`data` must be a numpy array")
        elif self.data.size and self.data.shape[0] != self.t.shape[0]:
            errors.append(f"first dimension of `data` ({self.data.shape[0]}) must match `t` ({self.t.shape[0]})")
        if self.frame.size and self.frame.shape[0] not in (1, self.t.shape[0]):
            errors.append(f"`frame` must have length 0, 1 or n_times; it has {self.frame.shape[0]}")
        if self.frameType not in range(len(FrameNames)):
            errors.append(f"frameType {self.frameType} is not a valid FrameType")
        if self.dataType not in range(len(DataNames)):
            errors.append(f"dataType {self.dataType} is not a valid DataType")
        if errors and assertions:
            raise ValueError("\n".join(errors))
        for e in errors:
            warnings.warn(e)
        return not errors

    # ------------------------------------------------------------------ bookkeeping
    @property
    def num(self):
        return self.__num

    def __str__(self):
        return f"{type(self).__name__}_{self.num}"

    def __repr__(self):
        return "\n".join(str(hh) for hh in self.history)

    def _append_history(self, hist, additional_depth=0):
        """Append reproducible command strings to the provenance log (scri/waveform_base.py:706-729)."""
        if not isinstance(hist, list):
            hist = [hist]
        self.history += [
            "# " * (self.__history_depth__ + additional_depth) + hist_line
            for hist_element in hist
            for hist_line in hist_element.split("\n")
        ]

    def _copy_kwargs(self):
        return dict(
            t=np.copy(self.t),
            frame=np.copy(self.frame),
            data=np.copy(self.data),
            history=self.history[:],
            version_hist=self.version_hist[:],
            frameType=self.frameType,
            dataType=self.dataType,
            r_is_scaled_out=self.r_is_scaled_out,
            m_is_scaled_out=self.m_is_scaled_out,
        )

    def copy(self):
        W = type(self)(**self._copy_kwargs(), constructor_statement=f"{self}.copy()")
        return W

    def interpolate(self, tprime):
        """Interpolate the frame and the data onto the new time steps (scri/waveform_base.py:949-967): the data through
        the not-a-knot cubic spline of every column (CubicSpline(t, data)(t') - on the GPU, scrib200_spline_remap with
        k = 1, alpha = 0), the frame through `squad`.  Returns a new object; only `t`, `frame` and `data` change."""
        from . import ops

        tprime = np.array(tprime, dtype=float)
        if tprime.ndim != 1:
            raise ValueError(f"New time array must have exactly 1 dimension; it has {tprime.ndim}.")
        kw = self._copy_kwargs()
        kw["t"] = tprime
        kw["frame"] = Q.squad(self.frame, self.t, tprime) if self.frame.shape[0] == self.t.shape[0] and self.frame.size else np.copy(self.frame)
        if self.data.size and tprime.size:
            kw["data"] = ops.spline_calculus(self.t, self.data.reshape(self.t.shape[0], -1), "evaluate", tprime=tprime).reshape(
                (tprime.shape[0],) + self.data.shape[1:])
        else:
            kw["data"] = np.empty((tprime.shape[0],) + self.data.shape[1:], dtype=self.data.dtype)
        return type(self)(**kw, constructor_statement=f"{self}.interpolate({np.array2string(tprime, threshold=6)})")

    def deepcopy(self):
        return _copy.deepcopy(self)

    # ------------------------------------------------------------------ simple properties
    @property
    def n_data_sets(self):
        return int(np.prod(self.data.shape[1:]))

    @property
    def n_times(self):
        return self.t.shape[0]

    @property
    def spin_weight(self):
        return SpinWeights[self.dataType]

    @property
    def conformal_weight(self):
        return ConformalWeights[self.dataType] - (RScaling[self.dataType] if self.r_is_scaled_out else 0)

    @property
    def gamma_weight(self):
        return (self.conformal_weight + self.spin_weight) / 2

    @property
    def r_scaling(self):
        return RScaling[self.dataType]

    @property
    def m_scaling(self):
        return MScaling[self.dataType]

    @property
    def frame_type_string(self):
        return FrameNames[self.frameType]

    @property
    def data_type_string(self):
        return DataNames[self.dataType]

    # ------------------------------------------------------------------ numerics (GPU)
    def norm(self, take_sqrt=False, indices=slice(None, None, None)):
        """L2 norm^2 of the data at each time (scri/waveform_base.py:535-551)."""
        from . import ops

        data = self.data if indices == slice(None, None, None) else self.data[indices]
        n = ops.norm(np.ascontiguousarray(data.reshape(data.shape[0], -1)))
        return np.sqrt(n) if take_sqrt else n

    def max_norm_index(self, skip_fraction_of_data=4):
        if skip_fraction_of_data == 0:
            return int(np.argmax(self.norm()))
        start = self.n_times // skip_fraction_of_data
        return int(np.argmax(self.norm(indices=slice(start, None))) + start)

    def max_norm_time(self, skip_fraction_of_data=4):
        return self.t[self.max_norm_index(skip_fraction_of_data=skip_fraction_of_data)]

    @property
    def data_dot(self):
        from . import ops

        return ops.spline_calculus(self.t, self.data, "derivative", 1)

    @property
    def data_ddot(self):
        from . import ops

        return ops.spline_calculus(self.t, self.data, "derivative", 2)

    @property
    def data_int(self):
        from . import ops

        return ops.spline_calculus(self.t, self.data, "antiderivative", 1)

    @property
    def data_iint(self):
        from . import ops

        return ops.spline_calculus(self.t, self.data, "antiderivative", 2)
