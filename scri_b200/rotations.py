"""Frame rotations of WaveformModes (mirrors scri/rotations.py), numerics on the GPU."""
import numpy as np

from . import _quaternion as Q
from . import ops
from .constants import Coorbital, Coprecessing, Corotating, Inertial
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
        frame = Q.qexp(log_frame)
    W.rotate_decomposition_basis(frame)
    W.frameType = Corotating
    W.__history_depth__ -= 1
    W._append_history(f"{W}.to_corotating_frame({R0}, {tolerance}, {z_alignment_region}, {return_omega}, {truncate_log_frame})")
    # the reference's four return shapes (scri/rotations.py:93-103); the RPXMB writers unpack `W, log_frame = ...`
    if return_omega:
        if truncate_log_frame:
            return (W, omega, log_frame)
        return (W, omega)
    if truncate_log_frame:
        return (W, log_frame)
    return W


@waveform_alterations
def to_coprecessing_frame(W, RoughDirection=np.array([0.0, 0.0, 1.0]), RoughDirectionIndex=None, transition_times=None):
    """Transform to a coprecessing frame via the dominant eigenvector of <LL> (scri/rotations.py:14-48).  With
    `transition_times` = (t_begin, t_end) the frame's angular velocity is switched off smoothly over that span and the
    frame re-integrated from t_begin on, so that it stops turning through the ringdown (scri/rotations.py:40-45)."""
    from .mode_calculations import LLDominantEigenvector, integrate_angular_velocity, minimal_rotation, rotor_angular_velocity
    from .sample_waveforms import transition_function

    if RoughDirectionIndex is None:
        RoughDirectionIndex = W.n_times // 8
    dpa = LLDominantEigenvector(W, RoughDirection=RoughDirection, RoughDirectionIndex=RoughDirectionIndex)
    # rotor taking z to dpa: sqrt(-dpa * z), dpa normalised as the reference does
    v = Q.qnormalized(np.concatenate([np.zeros((dpa.shape[0], 1)), dpa], axis=1))
    zq = np.array([0.0, 0.0, 0.0, 1.0])
    R = Q.qsqrt(-Q.qmul(v, zq))
    R = minimal_rotation(R, W.t, iterations=3)
    if transition_times is not None:
        i0, i1 = int(np.argmin(np.abs(W.t - transition_times[0]))), int(np.argmin(np.abs(W.t - transition_times[1])))
        transition = transition_function(W.t[i0:], W.t[i0], W.t[i1], y0=1.0, y1=0.0)
        omega = rotor_angular_velocity(R[i0:], W.t[i0:]) * transition[:, np.newaxis]
        slowingR = integrate_angular_velocity(W.t[i0:], omega, R0=R[i0])
        R = np.concatenate((R[:i0], slowingR))
    W.rotate_decomposition_basis(R)
    W.frameType = Coprecessing
    W.__history_depth__ -= 1
    W._append_history(f"{W}.to_coprecessing_frame({RoughDirection}, {RoughDirectionIndex}, {transition_times})")
    return W


def get_alignment_of_decomposition_frame_to_modes(w, t_fid, nHat_t_fid=(0.0, 1.0, 0.0, 0.0), ell_max=None):
    """Rotor R_eps that fixes the attitude of a corotating / coprecessing / coorbital frame at the fiducial time: the z axis
    of the decomposition frame goes onto the dominant eigenvector of <LL> (pointing along the angular velocity) and the
    phases of the (2, +-2) modes are made equal; the x axis ends up closer to `nHat_t_fid` than to its opposite
    (scri/rotations.py:114-225).  Returns a float rotor (w, x, y, z)."""
    import math

    from .mode_calculations import LLDominantEigenvector, angular_velocity

    if ell_max is None:
        ell_max = w.ell_max
    if w.frameType not in (Coprecessing, Coorbital, Corotating):
        raise ValueError(
            "get_alignment_of_decomposition_frame_to_modes only takes Waveforms in the coprecessing, coorbital, or corotating "
            f"frames.  This Waveform is in the '{w.frame_type_string}' frame."
        )
    if w.frame.shape[0] != w.n_times:
        raise ValueError(
            "get_alignment_of_decomposition_frame_to_modes requires full information about the Waveform's frame."
            f"This Waveform has {w.n_times} time steps, but only {w.frame.shape[0]} rotors in its frame."
        )
    if t_fid < w.t[0] or t_fid > w.t[-1]:
        raise ValueError(f"The requested alignment time t_fid={t_fid} is outside the range of times in this waveform ({w.t[0]}, {w.t[-1]}).")
    nHat = np.asarray(nHat_t_fid, dtype=float).ravel()[-3:]
    # direction of the angular velocity near t_fid, from an 11-sample window of the ell = 2 modes in the inertial frame
    i_t_fid = int((w.t <= t_fid).nonzero()[0][-1])
    if i_t_fid < w.t.size - 1:
        i_t_fid += 1
    i1 = max(0, i_t_fid - 5)
    i2 = w.t.size if i1 + 11 > w.t.size else i1 + 11
    region = w[i1:i2, 2].to_inertial_frame()
    om = angular_velocity(region)[i_t_fid - i1]
    omega_hat = np.concatenate([[0.0], om / np.linalg.norm(om)])
    R = w.frame[i_t_fid]
    omega_hat = Q.qmul(Q.qmul(Q.qinverse(R), omega_hat), R)        # components in this waveform's (rotating) frame
    instant = w[i1:i2].interpolate(np.array([t_fid]))
    R_f0 = instant.frame[0]
    V = LLDominantEigenvector(instant[:, : ell_max + 1])[0]
    V = V / np.linalg.norm(V)
    if np.dot(omega_hat[1:], V) < 0:
        V = -V
    zq = np.array([0.0, 0.0, 0.0, 1.0])
    R_V_f = Q.qsqrt(-Q.qmul(np.concatenate([[0.0], V]), zq))      # rotor taking z onto V
    instant.rotate_decomposition_basis(R_V_f)
    a22, a2m2 = instant.data[0, instant.index(2, 2)], instant.data[0, instant.index(2, -2)]
    phase = math.atan2(a22.imag, a22.real) - math.atan2(a2m2.imag, a2m2.real)
    R_eps = Q.qmul(R_V_f, Q.qexp_vec(np.array([0.0, 0.0, -phase / 8.0])))
    xq = np.array([0.0, 1.0, 0.0, 0.0])
    Rt = Q.qmul(R_f0, R_eps)
    if np.dot(nHat, Q.qmul(Q.qmul(Rt, xq), Q.qinverse(Rt))[1:]) < 0:
        R_eps = Q.qmul(R_eps, Q.qexp_vec(np.array([0.0, 0.0, math.pi / 2.0])))
    return R_eps


def align_decomposition_frame_to_modes(w, t_fid, nHat_t_fid=(0.0, 1.0, 0.0, 0.0), ell_max=None):
    """Fix the attitude of the corotating frame at `t_fid` by the constant rotor of
    `get_alignment_of_decomposition_frame_to_modes` (scri/rotations.py:228-265); rotates `w` in place and returns it."""
    R_eps = get_alignment_of_decomposition_frame_to_modes(w, t_fid, nHat_t_fid, ell_max)
    w._append_history(f"{w}.align_decomposition_frame_to_modes({t_fid}, {nHat_t_fid}, {ell_max})  # R_eps={R_eps}")
    return w.rotate_decomposition_basis(R_eps)
