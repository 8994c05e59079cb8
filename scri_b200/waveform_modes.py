"""WaveformModes: SWSH-mode waveform container with the reference's method surface.

Mirrors scri/waveform_modes.py (indexing :404-455, `transform` :705-719) and the operators that
scri/__init__.py:125-150 attaches to the class.  Data live in host numpy arrays as in the reference
(t float64 [n_times], data complex128 [n_times, n_modes], frame float [n_frames, 4]); every numerical
method stages them to the GPU, runs the sm_100a kernels and copies the result back.
"""
import warnings

import numpy as np

from . import _sf
from .constants import DataNames, SpinWeights, UnknownDataType
from .waveform_base import WaveformBase, waveform_alterations


class WaveformModes(WaveformBase):
    def _init_extra(self, args, kwargs):
        if len(args) == 0:
            self.__ell_min = kwargs.pop("ell_min", 0)
            self.__ell_max = kwargs.pop("ell_max", -1)
        else:
            self.__ell_min = args[0].ell_min
            self.__ell_max = args[0].ell_max
        self.__LM = _sf.LM_range(self.__ell_min, self.__ell_max) if self.__ell_max >= self.__ell_min else np.empty((0, 2), dtype=int)

    def ensure_validity(self, alter=True, assertions=False):
        ok = super().ensure_validity(alter=alter, assertions=assertions)
        errors = []
        if self.data.size:
            if self.data.dtype != np.dtype(complex):
                if alter:
                    self.data = np.asarray(self.data, dtype=complex)
                else:
                    errors.append("`data` must be complex")
            if self.data.ndim != 2:
                errors.append(f"`data` must be two-dimensional [time, mode]; it has shape {self.data.shape}")
            elif self.data.shape[1] != _sf.LM_total_size(self.ell_min, self.ell_max):
                errors.append(
                    f"second dimension of `data` ({self.data.shape[1]}) must equal the number of modes "
                    f"for ell_min={self.ell_min}, ell_max={self.ell_max} ({_sf.LM_total_size(self.ell_min, self.ell_max)})"
                )
        if self.dataType != UnknownDataType and self.ell_max >= 0 and self.ell_min < abs(SpinWeights[self.dataType]):
            pass  # the reference only warns about this in `ensure_validity`; modes below |s| are simply zero
        if errors and assertions:
            raise ValueError("\n".join(errors))
        for e in errors:
            warnings.warn(e)
        return ok and not errors

    def _copy_kwargs(self):
        kw = super()._copy_kwargs()
        kw.update(ell_min=self.ell_min, ell_max=self.ell_max)
        return kw

    # ------------------------------------------------------------------ (ell, m) layout
    @property
    def n_modes(self):
        return self.data.shape[1]

    @property
    def ell_min(self):
        return self.__ell_min

    @property
    def ell_max(self):
        return self.__ell_max

    @property
    def ells(self):
        return self.__ell_min, self.__ell_max

    @ells.setter
    def ells(self, new):
        self.__ell_min, self.__ell_max = new
        self.__LM = _sf.LM_range(self.__ell_min, self.__ell_max)

    @property
    def LM(self):
        """Array of [ell, m] pairs in storage order (scri/waveform_modes.py:404-418)."""
        return self.__LM

    def index(self, ell, m):
        """Flat index of mode (ell, m): ell(ell+1) - ell_min^2 + m (scri/waveform_modes.py:420-455)."""
        if ell < self.ell_min or ell > self.ell_max or abs(m) > ell:
            raise ValueError(f"(ell,m)=({ell},{m}) is not contained in this waveform with ell range [{self.ell_min},{self.ell_max}]")
        return _sf.LM_index(ell, m, self.ell_min)

    def indices(self, args):
        return [self.index(ell, m) for ell, m in args]

    # ------------------------------------------------------------------ BMS transformation
    def transform(self, **kwargs):
        """Apply a BMS transformation; returns a new WaveformModes (scri/waveform_modes.py:705-719).

        Keywords as in the reference: supertranslation, spacetime_translation, space_translation,
        time_translation, frame_rotation, boost_velocity, n_theta, n_phi, ell_max.
        """
        from .waveform_grid import WaveformGrid

        return WaveformGrid.transform(self, **kwargs)

    def to_grid(self, **kwargs):
        from .waveform_grid import WaveformGrid

        return WaveformGrid.from_modes(self, **kwargs)

    @classmethod
    def from_grid(cls, w_grid, ell_max):
        from .waveform_grid import WaveformGrid

        return WaveformGrid.to_modes(w_grid, ell_max)

    def __repr__(self):
        rep = super().__repr__()
        rep += f"\n# ell_min={self.ell_min}, ell_max={self.ell_max}"
        return rep


def _attach_operators():
    """The reference attaches its operators onto WaveformModes in scri/__init__.py:125-150."""
    from . import rotations

    WaveformModes.rotate_decomposition_basis = rotations.rotate_decomposition_basis
    WaveformModes.rotate_physical_system = rotations.rotate_physical_system
    WaveformModes.to_inertial_frame = rotations.to_inertial_frame
    WaveformModes.to_corotating_frame = rotations.to_corotating_frame
    WaveformModes.to_coprecessing_frame = rotations.to_coprecessing_frame


_attach_operators()
