"""Frame rotations of WaveformModes (mirrors scri/rotations.py), numerics on the GPU."""
import numpy as np

from . import _quaternion as Q
from . import ops
from .constants import Coprecessing, Corotating, Inertial
from .waveform_base import waveform_alterations


@waveform_alterations
def rotate_physical_system(W, R_phys):
    """Rotate the physical system = rotate the basis by the inverse (scri/rotations.py:268-281)."""
    R_phys = Q.as_float_quat(R_phys)
    W = rotate_decomposition_basis(W, Q.qconj(R_phys))
    W.__history_depth__ -= 1
    W._append_history(f"{W}.rotate_physical_system({R_phys})")
    return W


@waveform_alterations
def rotate_decomposition_basis(W, R_basis):
    """Rotate a waveform's decomposition basis in place and record it in `frame` (scri/rotations.py:284-343).

    `R_basis` is a single rotor (float [4], or np.quaternion when that package is available) or a series
    of n_times rotors ([n_times, 4]).  data <- data . D(R) per ell;  frame <- frame * R.
    """
    R_basis = Q.as_float_quat(R_basis)
    if R_basis.ndim == 2 and R_basis.shape[0] == 1:
        R_basis = R_basis[0]
    if R_basis.ndim == 2:
        if W.n_times != R_basis.shape[0]:
            raise ValueError(f"Input dimension mismatch.  (W.n_times={W.n_times}) != (len(R_basis)={R_basis.shape[0]})")
        ops.rotate_modes(W.data, R_basis, W.ell_min, W.ell_max)
        if W.frame.size:
            W.frame = Q.qmul(W.frame, R_basis)
        else:
            W.frame = np.copy(R_basis)
    elif R_basis.ndim == 1:
        ops.rotate_modes(W.data, R_basis, W.ell_min, W.ell_max)
        if W.frame.size:
            W.frame = Q.qmul(W.frame, R_basis)
        else:
            W.frame = np.array([R_basis])
    else:
        raise ValueError(f"Input dimension mismatch.  R_basis.shape={R_basis.shape}")
    opts = np.get_printoptions()
    np.set_printoptions(threshold=6)
    W.__history_depth__ -= 1
    W._append_history(f"{W}.rotate_decomposition_basis({R_basis})")
    np.set_printoptions(**opts)
    return W


@waveform_alterations
def to_inertial_frame(W):
    """Undo the recorded frame rotation (scri/rotations.py:106-111)."""
    W.rotate_decomposition_basis(Q.qconj(W.frame))
    W.frameType = Inertial
    W.__history_depth__ -= 1
    W._append_history(f"{W}.to_inertial_frame()")
    return W


@waveform_alterations
def to_corotating_frame(W, R0=(1.0, 0.0, 0.0, 0.0), tolerance=1e-12, z_alignment_region=None, return_omega=False, truncate_log_frame=False):
    """Transform to the corotating frame (scri/rotations.py:51-103)."""
    from .mode_calculations import corotating_frame

    frame, omega = corotating_frame(W, R0=R0, tolerance=tolerance, z_alignment_region=z_alignment_region, return_omega=True)
    if truncate_log_frame:
        log_frame = Q.qlog(frame)
        power_of_2 = 2 ** int(-np.floor(np.log2(2 * tolerance)))
        log_frame = np.round(log_frame * power_of_2) / power_of_2
        frame = Q.qexp_vec(log_frame[..., 1:])
    W.rotate_decomposition_basis(frame)
    W.frameType = Corotating
    W.__history_depth__ -= 1
    W._append_history(f"{W}.to_corotating_frame({R0}, {tolerance}, {z_alignment_region}, {return_omega}, {truncate_log_frame})")
    if return_omega:
        return (W, omega)
    return W


@waveform_alterations
def to_coprecessing_frame(W, RoughDirection=np.array([0.0, 0.0, 1.0]), RoughDirectionIndex=None):
    """Transform to a coprecessing frame via the dominant eigenvector of <LL> (scri/rotations.py:14-48)."""
    from .mode_calculations import LLDominantEigenvector, minimal_rotation

    if RoughDirectionIndex is None:
        RoughDirectionIndex = W.n_times // 8
    dpa = LLDominantEigenvector(W, RoughDirection=RoughDirection, RoughDirectionIndex=RoughDirectionIndex)
    # rotor taking z to dpa: sqrt(-dpa * z)
    v = np.concatenate([np.zeros((dpa.shape[0], 1)), dpa], axis=1)
    zq = np.array([0.0, 0.0, 0.0, 1.0])
    R = Q.qsqrt(-Q.qmul(v, zq))
    R = minimal_rotation(R, W.t, iterations=3)
    W.rotate_decomposition_basis(R)
    W.frameType = Coprecessing
    W.__history_depth__ -= 1
    W._append_history(f"{W}.to_coprecessing_frame({RoughDirection}, {RoughDirectionIndex})")
    return W
