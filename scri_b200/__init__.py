"""scri_b200: B200-native (sm_100a) implementation of scri's waveform-transformation hot path.

Drop-in for the numba/spinsfast path behind scri's own Python API: `WaveformModes`, `WaveformGrid`
(`to_grid`, `from_grid`, `transform`, `rotate_decomposition_basis`, `to_corotating_frame`, ...),
`mode_calculations` and `flux` entry points.  Python host code calls a thin C-ABI CUDA library
(include/scrib200.h) through PyTorch tensors; there is no CPU fallback.
"""
from .constants import (  # noqa: F401
    ConformalWeights, Coorbital, Coprecessing, Corotating, DataNames, DataType, FrameNames, FrameType, Inertial,
    MScaling, RScaling, SpinWeights, UnknownDataType, UnknownFrameType,
    h, hdot, news, psi0, psi1, psi2, psi3, psi4, psim, psin, sigma,
)
from .waveform_base import WaveformBase  # noqa: F401
from .waveform_modes import WaveformModes
from .waveform_grid import WaveformGrid  # noqa: F401
from .mode_calculations import (  # noqa: F401
    LdtVector, LVector, LLMatrix, LLComparisonMatrix, LLDominantEigenvector, angular_velocity, corotating_frame,
)
from .flux import energy_flux, momentum_flux, angular_momentum_flux, boost_flux, poincare_fluxes  # noqa: F401
from .rotations import (  # noqa: F401
    align_decomposition_frame_to_modes, get_alignment_of_decomposition_frame_to_modes, rotate_decomposition_basis, rotate_physical_system,
    to_coprecessing_frame, to_corotating_frame, to_inertial_frame,
)
from . import sample_waveforms  # noqa: F401
from .asymptotic_bondi_data import AsymptoticBondiData, ModesTimeSeries  # noqa: F401

# operators attached to the class, as scri/__init__.py:125-150 does
WaveformModes.LdtVector = LdtVector
WaveformModes.LVector = LVector
WaveformModes.LLMatrix = LLMatrix
WaveformModes.LLComparisonMatrix = LLComparisonMatrix
WaveformModes.LLDominantEigenvector = LLDominantEigenvector
WaveformModes.angular_velocity = angular_velocity
WaveformModes.energy_flux = energy_flux
WaveformModes.momentum_flux = momentum_flux
WaveformModes.angular_momentum_flux = angular_momentum_flux
WaveformModes.boost_flux = boost_flux
WaveformModes.poincare_fluxes = poincare_fluxes

__version__ = "0.1.0"
