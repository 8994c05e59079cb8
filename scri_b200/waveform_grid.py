"""WaveformGrid: waveform sampled on the (theta, phi) grid, and the BMS transformation entry points.

Mirrors scri/waveform_grid.py: `from_modes` (:331-613), `to_modes` (:274-329), `transform` (:615-630),
with the numerics executed on the GPU by scri_b200.plan.TransformPlan.
"""
import numbers
import os
import pprint
import warnings

import numpy as np

from . import _lib, _sf, ops
from .constants import Inertial, SpinWeights
from .plan import TransformPlan, _trace, cached_transform_plan
from .waveform_base import WaveformBase
from .waveform_modes import WaveformModes

# relative sizes of the slabs the modes are uploaded in (small first: the kernels start early; small last: a short tail).
# SCRIB200_SLAB_WEIGHTS is a developer knob for measuring other cuts.
_SLAB_WEIGHTS = tuple(float(x) for x in os.environ.get("SCRIB200_SLAB_WEIGHTS", "1,2,3,4,4,4,3,2,1").split(","))




class WaveformGrid(WaveformBase):
    def _init_extra(self, args, kwargs):
        if len(args) == 0:
            self.__n_theta = kwargs.pop("n_theta", 0)
            self.__n_phi = kwargs.pop("n_phi", 0)
        else:
            self.__n_theta = args[0].n_theta
            self.__n_phi = args[0].n_phi

    def ensure_validity(self, alter=True, assertions=False):
        ok = super().ensure_validity(alter=alter, assertions=assertions)
        errors = []
        if self.data.size and self.data.shape[1] != self.n_theta * self.n_phi:
            errors.append(
                f"second dimension of `data` ({self.data.shape[1]}) must equal n_theta*n_phi = {self.n_theta * self.n_phi}"
            )
        if errors and assertions:
            raise ValueError("\n".join(errors))
        for e in errors:
            warnings.warn(e)
        return ok and not errors

    def _copy_kwargs(self):
        kw = super()._copy_kwargs()
        kw.update(n_theta=self.n_theta, n_phi=self.n_phi)
        return kw

    @property
    def n_theta(self):
        return self.__n_theta

    @property
    def n_phi(self):
        return self.__n_phi

    def to_modes(self, ell_max=None, ell_min=None):
        """SWSH analysis of every time step (scri/waveform_grid.py:274-329)."""
        s = SpinWeights[self.dataType]
        if ell_max is None:
            ell_max = int((max(self.n_theta, self.n_phi) - 1) // 2)
        if ell_min is None:
            ell_min = abs(s)
        if not isinstance(ell_max, numbers.Integral) or ell_max < 0:
            raise ValueError(f"Input `ell_max` should be a nonnegative integer; got `{ell_max}`.")
        if not isinstance(ell_min, numbers.Integral) or ell_min < 0 or ell_min > ell_max:
            raise ValueError(f"Input `ell_min` should be an integer between 0 and {ell_max}; got `{ell_min}`.")
        if self.data.ndim != 2:
            raise ValueError("scri_b200 supports two-dimensional grid data [time, n_theta*n_phi] only")
        new_data = ops.map2salm(self.data, s, int(ell_max), self.n_theta, self.n_phi, ell_min=int(ell_min))
        return WaveformModes(
            t=self.t,
            data=new_data,
            history=self.history,
            ell_min=ell_min,
            ell_max=ell_max,
            frameType=self.frameType,
            dataType=self.dataType,
            r_is_scaled_out=self.r_is_scaled_out,
            m_is_scaled_out=self.m_is_scaled_out,
            constructor_statement=f"{self}.to_modes({ell_max})",
        )

    @classmethod
    def from_modes(cls, w_modes, **kwargs):
        """Evaluate modes on the (BMS-transformed) grid (scri/waveform_grid.py:331-613)."""
        if not isinstance(w_modes, WaveformModes):
            raise TypeError(
                f"\nInput waveform object must be an instance of `WaveformModes`; this is of type `{type(w_modes).__name__}`"
            )
        if w_modes.frameType != Inertial:
            raise ValueError(
                f"\nInput waveform object must be in an inertial frame; this is in a frame of type `{w_modes.frame_type_string}`"
            )
        if w_modes.data.ndim != 2:
            raise ValueError("scri_b200 supports two-dimensional mode data [time, mode] only")
        original_kwargs = kwargs.copy()
        a_fut = ops.to_device_async(w_modes.data, np.complex128)   # streams in while the plan is built
        try:
            plan = cached_transform_plan(
                w_modes.ell_min, w_modes.ell_max, w_modes.dataType, r_is_scaled_out=w_modes.r_is_scaled_out, **kwargs
            )
        finally:
            a_d = a_fut.result()
        t_d = ops.to_device(w_modes.t, np.float64)
        uprm, grid = plan.run(t_d, a_d, return_grid=True)
        g = cls(
            t=ops.to_host(uprm),
            data=ops.to_host(grid),
            history=w_modes.history,
            n_theta=plan.n_theta,
            n_phi=plan.n_phi,
            frameType=w_modes.frameType,
            dataType=w_modes.dataType,
            r_is_scaled_out=w_modes.r_is_scaled_out,
            m_is_scaled_out=w_modes.m_is_scaled_out,
            constructor_statement=f"{cls.__name__}.from_modes({w_modes}, **{original_kwargs})",
        )
        if plan.leftover_kwargs:
            warnings.warn("\nUnused kwargs passed to this function:\n{}".format(pprint.pformat(plan.leftover_kwargs, width=1)))
        return g

    @classmethod
    def transform(cls, w_modes, **kwargs):
        """from_modes followed by to_modes, the grid staying on the device (scri/waveform_grid.py:615-630)."""
        if not isinstance(w_modes, WaveformModes):
            raise TypeError(f"Expected WaveformModes object in argument 1; got `{type(w_modes).__name__}` instead.")
        ell_max = kwargs.pop("ell_max", w_modes.ell_max)
        if w_modes.frameType != Inertial:
            raise ValueError(
                f"\nInput waveform object must be in an inertial frame; this is in a frame of type `{w_modes.frame_type_string}`"
            )
        original_kwargs = kwargs.copy()
        # the modes stream in slab by slab on a copy stream while the plan is built; each slab is synthesized as it lands
        big = w_modes.data.nbytes >= (8 << 20)
        t_d = ops.to_device_nowait(w_modes.t, np.float64)   # before the modes: a copy queued behind them would wait for all of them
        if big:
            a_d, slabs, a_fut = ops.to_device_slabs(w_modes.data, np.complex128, weights=_SLAB_WEIGHTS)
        else:
            a_d, slabs, a_fut = ops.to_device(w_modes.data, np.complex128), None, None
        try:
            plan = cached_transform_plan(
                w_modes.ell_min, w_modes.ell_max, w_modes.dataType, r_is_scaled_out=w_modes.r_is_scaled_out,
                out_ell_max=ell_max, **kwargs,
            )
        except BaseException:
            if a_fut is not None:
                a_fut.result()                   # let the copy thread finish with w_modes.data before unwinding
            raise
        if plan.mix and slabs is not None:       # the Weyl mixing reads whole fields: wait for the transfer
            a_fut.result()
            _lib.require_cuda().cuda.current_stream().wait_event(slabs[-1][2])
            slabs = None
        # the provenance string is formatted and the result object built (time-axis checks included) while the pipeline runs:
        # neither before the first launch nor after the last byte.  `early` receives the arrays the result will arrive in
        note = {}

        def wrap(u, m, statement):
            return WaveformModes(
                t=u,
                data=m,
                history=w_modes.history,
                ell_min=plan.out_ell_min,
                ell_max=plan.out_ell_max,
                frameType=w_modes.frameType,
                dataType=w_modes.dataType,
                r_is_scaled_out=w_modes.r_is_scaled_out,
                m_is_scaled_out=w_modes.m_is_scaled_out,
                constructor_statement=statement,
            )

        def early(u=None, m=None):
            note["statement"] = f"WaveformGrid.from_modes({w_modes}, **{original_kwargs}).to_modes({ell_max})"
            if m is not None:
                note["arrays"] = (u, m)
                note["result"] = wrap(u, m, note["statement"])

        # modes land in pinned host memory slab by slab, the first output slabs while the last input slabs are still in flight
        uprm, modes = plan.run(t_d, a_d, slabs=slabs, host_slabs=4, t_host=np.asarray(w_modes.t, dtype=float), on_queued=early)
        _trace("plan.run returned")
        if a_fut is not None:
            a_fut.result()                       # surfaces a failed copy
        if plan.leftover_kwargs:
            warnings.warn("\nUnused kwargs passed to this function:\n{}".format(pprint.pformat(plan.leftover_kwargs, width=1)))
        if "result" in note and note["arrays"][0] is uprm and note["arrays"][1] is modes:
            return note["result"]                # (a repeated pipeline - time axis changed in place - returns other arrays)
        if "statement" not in note:
            early()
        return wrap(ops.to_host(uprm), ops.to_host(modes), note["statement"])

    def __repr__(self):
        rep = super().__repr__()
        rep += f"\n# n_theta={self.n_theta}, n_phi={self.n_phi}"
        return rep
