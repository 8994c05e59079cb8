"""CPU restatement of scri's hot-path flow (reference: /root/reference/scri, v2024.0.13).

TEST INFRASTRUCTURE (see oracle/__init__.py) - never imported by scri_b200/.

Each function cites the reference lines it follows.  Waveforms are plain `Modes` records
(t, data, ell_min, ell_max, dataType, ...) rather than the reference's class hierarchy; the
arithmetic (order of operations, scipy spline classes, numba loop bodies) follows the reference.
Quaternions are float [...,4] arrays (oracle.quat) since numpy-quaternion is absent.
"""
import math
import warnings
from dataclasses import dataclass, field, replace
from functools import lru_cache

import numpy as np
import numba
from scipy import interpolate
from scipy.interpolate import CubicSpline

from . import quat, sf, spinsfast

jit = numba.njit(cache=False)

# scri/__init__.py:78-86
FrameType = [UnknownFrameType, Inertial, Coprecessing, Coorbital, Corotating] = range(5)
DataType = [UnknownDataType, psi0, psi1, psi2, psi3, psi4, sigma, h, hdot, news, psin, psim] = range(12)
DataNames = ["UnknownDataType", "Psi0", "Psi1", "Psi2", "Psi3", "Psi4", "sigma", "h", "hdot", "news", "psin", "PsiM"]
_big = 2**63 - 1
SpinWeights = [_big, 2, 1, 0, -1, -2, 2, -2, -2, -2, _big, 0]
ConformalWeights = [_big, 2, 1, 0, -1, -2, 1, 0, -1, -1, -3, 0]
RScaling = [_big, 5, 4, 3, 2, 1, 2, 1, 1, 1, 0, 0]
MScaling = [_big, 2, 2, 2, 2, 2, 0, 0, 1, 1, 2, 1]


@dataclass
class Modes:
    t: np.ndarray
    data: np.ndarray
    ell_min: int = 2
    ell_max: int = 8
    dataType: int = h
    frameType: int = Inertial
    r_is_scaled_out: bool = True
    m_is_scaled_out: bool = True
    frame: np.ndarray = field(default_factory=lambda: np.zeros((0, 4)))

    @property
    def spin_weight(self):
        return SpinWeights[self.dataType]

    @property
    def conformal_weight(self):
        # scri/waveform_base.py:444-446
        return ConformalWeights[self.dataType] - (RScaling[self.dataType] if self.r_is_scaled_out else 0)

    @property
    def n_times(self):
        return self.t.shape[0]

    @property
    def LM(self):
        return sf.LM_range(self.ell_min, self.ell_max)

    def copy(self):
        return replace(self, t=self.t.copy(), data=self.data.copy(), frame=np.array(self.frame, copy=True))


@dataclass
class Grid:
    t: np.ndarray
    data: np.ndarray  # [n_times, n_theta*n_phi]
    n_theta: int
    n_phi: int
    dataType: int = h
    frameType: int = Inertial
    r_is_scaled_out: bool = True
    m_is_scaled_out: bool = True


# ----------------------------------------------------------------------------- waveform_grid.py:20-190
def process_transformation_kwargs(ell_max, **kwargs):
    supertranslation = np.zeros((4,), dtype=complex)
    ell_max_supertranslation = 1
    if "supertranslation" in kwargs:
        supertranslation = np.array(kwargs.pop("supertranslation"), dtype=complex)
        if supertranslation.size <= 4:
            supertranslation = np.pad(supertranslation, (0, 4 - supertranslation.size), "constant", constant_values=(0.0,))
        ell_max_supertranslation = int(np.sqrt(len(supertranslation))) - 1
        if (ell_max_supertranslation + 1) ** 2 != len(supertranslation):
            raise ValueError("supertranslation length must be a perfect square; it is {}".format(len(supertranslation)))
        for ell in range(ell_max_supertranslation + 1):
            for m in range(ell + 1):
                i_pos = sf.LM_index(ell, m, 0)
                i_neg = sf.LM_index(ell, -m, 0)
                a = supertranslation[i_pos]
                b = supertranslation[i_neg]
                if abs(a - (-1.0) ** m * b.conjugate()) > 3e-16 + 1e-15 * abs(b):
                    raise ValueError("Will result in an imaginary supertranslation.")
    spacetime_translation = np.zeros((4,), dtype=float)
    spacetime_translation[0] = sf.constant_from_ell_0_mode(supertranslation[0]).real
    spacetime_translation[1:4] = -sf.vector_from_ell_1_modes(supertranslation[1:4]).real
    if "spacetime_translation" in kwargs:
        st_trans = np.array(kwargs.pop("spacetime_translation"), dtype=float)
        if st_trans.shape != (4,):
            raise TypeError("spacetime_translation should be a float array of shape (4,)")
        spacetime_translation = st_trans[:]
        supertranslation[0] = sf.constant_as_ell_0_mode(spacetime_translation[0])
        supertranslation[1:4] = sf.vector_as_ell_1_modes(-spacetime_translation[1:4])
    if "space_translation" in kwargs:
        s_trans = np.array(kwargs.pop("space_translation"), dtype=float)
        if s_trans.shape != (3,):
            raise TypeError("space_translation should be an array of floats of shape (3,)")
        spacetime_translation[1:4] = s_trans[:]
        supertranslation[1:4] = sf.vector_as_ell_1_modes(-spacetime_translation[1:4])
    if "time_translation" in kwargs:
        t_trans = kwargs.pop("time_translation")
        if not isinstance(t_trans, float):
            raise TypeError("time_translation should be a single float")
        spacetime_translation[0] = t_trans
        supertranslation[0] = sf.constant_as_ell_0_mode(spacetime_translation[0])

    w_ell_max = ell_max
    ell_max = w_ell_max + ell_max_supertranslation
    n_theta = kwargs.pop("n_theta", 2 * ell_max + 1)
    n_phi = kwargs.pop("n_phi", 2 * ell_max + 1)
    if n_theta < 2 * ell_max + 1 and abs(supertranslation[1:]).max() > 0.0:
        warnings.warn(f"n_theta={n_theta} is small; because of the supertranslation, it will lose accuracy")
    if n_theta < 2 * w_ell_max + 1:
        raise ValueError(f"n_theta={n_theta} is too small; must be at least 2*ell+1={2 * w_ell_max + 1}")
    if n_phi < 2 * ell_max + 1 and abs(supertranslation[1:]).max() > 0.0:
        warnings.warn(f"n_phi={n_phi} is small; because of the supertranslation, it will lose accuracy")
    if n_phi < 2 * w_ell_max + 1:
        raise ValueError(f"n_phi={n_phi} is too small; must be at least 2*ell+1={2 * w_ell_max + 1}")

    frame_rotation = np.array(kwargs.pop("frame_rotation", [1, 0, 0, 0]), dtype=float)
    if quat.absq(frame_rotation) < 3e-16:
        raise ValueError(f"frame_rotation={frame_rotation} should be a unit quaternion")
    frame_rotation = quat.normalized(frame_rotation)

    boost_velocity = np.array(kwargs.pop("boost_velocity", [0.0] * 3), dtype=float)
    beta = np.linalg.norm(boost_velocity)
    if boost_velocity.shape != (3,) or beta >= 1.0:
        raise ValueError("boost_velocity should be a 3-vector with magnitude strictly less than 1.0.")
    gamma = 1 / math.sqrt(1 - beta**2)
    varphi = math.atanh(beta)

    thetaprm = np.linspace(0.0, np.pi, num=n_theta, endpoint=True)
    phiprm = np.linspace(0.0, 2 * np.pi, num=n_phi, endpoint=False)

    if beta > 3e-14:
        vhat = boost_velocity / beta

        def Bprm_j_k(thetaprm_, phiprm_):
            rprm = np.array(
                [math.cos(phiprm_) * math.sin(thetaprm_), math.sin(phiprm_) * math.sin(thetaprm_), math.cos(thetaprm_)]
            )
            Thetaprm = math.acos(np.dot(vhat, rprm))
            Theta = 2 * math.atan(math.exp(-varphi) * math.tan(Thetaprm / 2.0))
            c = np.cross(rprm, vhat)
            cq = np.array([0.0, c[0], c[1], c[2]])
            if quat.absq(cq) > 1e-200:
                return quat.exp(quat.normalized(cq) * (Thetaprm - Theta) / 2)
            else:
                return quat.one.copy()

    else:

        def Bprm_j_k(thetaprm_, phiprm_):
            return quat.one.copy()

    R_j_k = np.empty((n_theta, n_phi, 4), dtype=float)
    for j in range(n_theta):
        for k in range(n_phi):
            rotated = quat.mul(frame_rotation, quat.from_spherical_coords(thetaprm[j], phiprm[k]))
            th, ph = quat.as_spherical_coords(rotated)
            R_j_k[j, k] = quat.mul(Bprm_j_k(th, ph), rotated)

    return (
        supertranslation,
        ell_max_supertranslation,
        ell_max,
        n_theta,
        n_phi,
        boost_velocity,
        beta,
        gamma,
        varphi,
        R_j_k,
        kwargs,
    )


# ----------------------------------------------------------------------------- waveform_grid.py:331-613
def from_modes(w_modes, return_intermediates=False, **kwargs):
    if w_modes.frameType != Inertial:
        raise ValueError("Input waveform object must be in an inertial frame")
    (
        supertranslation,
        ell_max_supertranslation,
        ell_max,
        n_theta,
        n_phi,
        boost_velocity,
        beta,
        gamma,
        varphi,
        R_j_k,
        kwargs,
    ) = process_transformation_kwargs(w_modes.ell_max, **kwargs)

    SWSH_j_k = sf.SWSH_grid(R_j_k, w_modes.spin_weight, ell_max)
    SH_j_k = sf.SWSH_grid(R_j_k, 0, ell_max_supertranslation)
    r_j_k = quat.rotate_vector(R_j_k.reshape(-1, 4), np.array([0.0, 0.0, 1.0])).T  # [3, G]
    kconformal_j_k = 1.0 / (gamma * (1 - np.dot(boost_velocity, r_j_k).reshape(R_j_k.shape[:2])))
    alphasupertranslation_j_k = np.tensordot(supertranslation, SH_j_k, axes=([0], [2])).real
    fprm_i_j_k = np.tensordot(
        w_modes.data,
        SWSH_j_k[:, :, sf.LM_index(w_modes.ell_min, -w_modes.ell_min, 0) : sf.LM_index(w_modes.ell_max, w_modes.ell_max, 0) + 1],
        axes=([1], [2]),
    )
    if beta != 0 or (supertranslation[1:] != 0).any():
        if w_modes.dataType == h:
            supertranslation_deriv = 2 * sf.ethbar_GHP(sf.ethbar_GHP(supertranslation, 0, 0), -1, 0)
            vals = np.tensordot(
                supertranslation_deriv,
                SWSH_j_k[:, :, : sf.LM_index(ell_max_supertranslation, ell_max_supertranslation, 0) + 1],
                axes=([0], [2]),
            )
            fprm_i_j_k -= vals[np.newaxis, :, :]
        elif w_modes.dataType == sigma:
            supertranslation_deriv = sf.eth_GHP(sf.eth_GHP(supertranslation, 0, 0), 1, 0)
            vals = np.tensordot(
                supertranslation_deriv,
                SWSH_j_k[:, :, : sf.LM_index(ell_max_supertranslation, ell_max_supertranslation, 0) + 1],
                axes=([0], [2]),
            )
            fprm_i_j_k -= vals[np.newaxis, :, :]
        elif w_modes.dataType in [psi0, psi1, psi2, psi3]:
            from scipy.special import comb

            eth_alpha_j_k = np.tensordot(
                1 / np.sqrt(2) * sf.eth_GHP(supertranslation, spin_weight=0),
                sf.SWSH_grid(R_j_k, 1, ell_max_supertranslation),
                axes=([0], [2]),
            )
            v_dot_rhat = np.insert(sf.vector_as_ell_1_modes(boost_velocity), 0, 0.0)
            eth_v_dot_rhat_j_k = np.tensordot(1 / np.sqrt(2) * v_dot_rhat, sf.SWSH_grid(R_j_k, 1, 1), axes=([0], [2]))
            eth_uprm_over_k = (
                w_modes.t[:, np.newaxis, np.newaxis] - alphasupertranslation_j_k[np.newaxis, :, :]
            ) * gamma * kconformal_j_k[np.newaxis, :, :] * eth_v_dot_rhat_j_k[np.newaxis, :, :] - eth_alpha_j_k[np.newaxis, :, :]
            for DT in range(w_modes.dataType + 1, psi4 + 1):
                try:
                    w_temp = kwargs.pop("psi{}_modes".format(DataNames[DT][-1]))
                except KeyError:
                    raise ValueError(
                        "A BMS transformation of {} requires information from {}, which has not been supplied.".format(
                            DataNames[w_modes.dataType], DataNames[DT]
                        )
                    )
                SWSH_temp = sf.SWSH_grid(R_j_k, w_temp.spin_weight, w_temp.ell_max)
                f_i_j_k = np.tensordot(
                    w_temp.data,
                    SWSH_temp[:, :, sf.LM_index(w_temp.ell_min, -w_temp.ell_min, 0) : sf.LM_index(w_temp.ell_max, w_temp.ell_max, 0) + 1],
                    axes=([1], [2]),
                )
                fprm_i_j_k += comb(5 - w_modes.dataType, 5 - DT) * f_i_j_k * eth_uprm_over_k ** (DT - w_modes.dataType)
        elif w_modes.dataType not in [psi4, hdot, news]:
            warnings.warn("No BMS transformation is implemented for this dataType; proceeding as Psi4")

    fprm_i_j_k *= (kconformal_j_k**w_modes.conformal_weight)[np.newaxis, :, :]
    synthesized = fprm_i_j_k.copy() if return_intermediates else None

    time_translation = sf.constant_from_ell_0_mode(supertranslation[0]).real
    uprm_i = (1 / gamma) * (w_modes.t - time_translation)
    uprm_min = (kconformal_j_k * (w_modes.t[0] - alphasupertranslation_j_k)).max()
    uprm_max = (kconformal_j_k * (w_modes.t[-1] - alphasupertranslation_j_k)).min()
    uprm_iprm = uprm_i[(uprm_i >= uprm_min) & (uprm_i <= uprm_max)]

    for j in range(n_theta):
        for k in range(n_phi):
            uprm_i_j_k = kconformal_j_k[j, k] * (w_modes.t - alphasupertranslation_j_k[j, k])
            re = interpolate.InterpolatedUnivariateSpline(uprm_i_j_k, fprm_i_j_k[:, j, k].real)
            im = interpolate.InterpolatedUnivariateSpline(uprm_i_j_k, fprm_i_j_k[:, j, k].imag)
            fprm_i_j_k[: len(uprm_iprm), j, k] = re(uprm_iprm) + 1j * im(uprm_iprm)

    fprm_iprm_j_k = np.delete(fprm_i_j_k, np.s_[len(uprm_iprm) :], 0)
    fprm_iprm_j_k = fprm_iprm_j_k.reshape((fprm_iprm_j_k.shape[0], n_theta * n_phi))

    g = Grid(
        t=uprm_iprm,
        data=fprm_iprm_j_k,
        n_theta=n_theta,
        n_phi=n_phi,
        dataType=w_modes.dataType,
        frameType=w_modes.frameType,
        r_is_scaled_out=w_modes.r_is_scaled_out,
        m_is_scaled_out=w_modes.m_is_scaled_out,
    )
    if kwargs:
        warnings.warn("Unused kwargs passed to this function: {}".format(kwargs))
    if return_intermediates:
        return g, dict(
            synthesized=synthesized, kconformal=kconformal_j_k, alpha=alphasupertranslation_j_k, R_j_k=R_j_k, SWSH_j_k=SWSH_j_k
        )
    return g


# ----------------------------------------------------------------------------- waveform_grid.py:274-329
def to_modes(g, ell_max=None, ell_min=None):
    s = SpinWeights[g.dataType]
    if ell_max is None:
        ell_max = int((max(g.n_theta, g.n_phi) - 1) // 2)
    if ell_min is None:
        ell_min = abs(s)
    old = g.data.reshape((g.t.shape[0], g.n_theta, g.n_phi))
    # the reference loops over time calling spinsfast.map2salm per step; map2salm is batched over
    # leading axes, so one call is the same arithmetic
    new = spinsfast.map2salm(old, s, ell_max)[:, sf.LM_index(ell_min, -ell_min, 0) :]
    return Modes(
        t=g.t,
        data=new,
        ell_min=ell_min,
        ell_max=ell_max,
        dataType=g.dataType,
        frameType=g.frameType,
        r_is_scaled_out=g.r_is_scaled_out,
        m_is_scaled_out=g.m_is_scaled_out,
    )


def transform(w_modes, **kwargs):
    """waveform_grid.py:615-630 / waveform_modes.py:705-719"""
    ell_max = kwargs.pop("ell_max", w_modes.ell_max)
    return to_modes(from_modes(w_modes, **kwargs), ell_max)


# ----------------------------------------------------------------------------- rotations.py:346-392
@jit
def _rotate_by_constant(data, ell_min, ell_max, D, tmp):
    for i_t in range(data.shape[0]):
        for ell in range(ell_min, ell_max + 1):
            i_data = ell**2 - ell_min**2
            i_D = ((4 * ell**2 - 1) * ell - (4 * ell_min**2 - 1) * ell_min) // 3
            for i_m in range(2 * ell + 1):
                tmp[i_m] = 0j
            for i_mp in range(2 * ell + 1):
                for i_m in range(2 * ell + 1):
                    tmp[i_m] += data[i_t, i_data + i_mp] * D[i_D + (2 * ell + 1) * i_mp + i_m]
            for i_m in range(2 * ell + 1):
                data[i_t, i_data + i_m] = tmp[i_m]


@jit
def _rotate_by_series(data, Dall, ell_min, ell_max):
    # rotations.py:370-392 with the per-step sf._Wigner_D_matrices call hoisted into `Dall[i_t]`
    for i_t in range(data.shape[0]):
        D = Dall[i_t]
        for ell in range(ell_min, ell_max + 1):
            i_data = ell**2 - ell_min**2
            i_D = ((4 * ell**2 - 1) * ell - (4 * ell_min**2 - 1) * ell_min) // 3
            for i_m in range(2 * ell + 1):
                new_data_mp = 0j
                for i_mp in range(2 * ell + 1):
                    new_data_mp += data[i_t, i_data + i_mp] * D[i_D + i_m + (2 * ell + 1) * i_mp]
                D[i_D + i_m] = new_data_mp
            for i_m in range(2 * ell + 1):
                data[i_t, i_data + i_m] = D[i_D + i_m]


def rotate_decomposition_basis(W, R_basis):
    """rotations.py:284-343 (in place; frame <- frame * R)."""
    R_basis = np.asarray(R_basis, dtype=float)
    if R_basis.ndim == 2 and R_basis.shape[0] == 1:
        R_basis = R_basis[0]
    if R_basis.ndim == 2:
        if W.n_times != len(R_basis):
            raise ValueError("Input dimension mismatch.")
        sp = quat.as_spinor_array(R_basis)
        # chunk to bound the D workspace
        step = 4096
        for a in range(0, W.n_times, step):
            Dall = sf.Wigner_D_matrices(sp[a : a + step, 0], sp[a : a + step, 1], W.ell_min, W.ell_max)
            _rotate_by_series(W.data[a : a + step], Dall, W.ell_min, W.ell_max)
        if W.frame.size:
            W.frame = quat.mul(W.frame, R_basis)
        else:
            W.frame = np.copy(R_basis)
    else:
        sp = quat.as_spinor_array(R_basis)
        D = sf.Wigner_D_matrices(sp[0], sp[1], W.ell_min, W.ell_max)
        tmp = np.empty((2 * W.ell_max + 1,), dtype=complex)
        _rotate_by_constant(W.data, W.ell_min, W.ell_max, D, tmp)
        if W.frame.size:
            W.frame = quat.mul(W.frame, R_basis)
        else:
            W.frame = np.array([R_basis])
    return W


def to_inertial_frame(W):
    """rotations.py:106-111"""
    W = rotate_decomposition_basis(W, quat.conj(W.frame))
    W.frameType = Inertial
    return W


# ----------------------------------------------------------------------------- waveform_base.py:689-703
def data_dot(W):
    return CubicSpline(W.t, W.data.view(float), axis=0).derivative()(W.t).view(complex) if False else CubicSpline(
        W.t, W.data
    ).derivative()(W.t)


def data_ddot(W):
    return CubicSpline(W.t, W.data).derivative(2)(W.t)


def data_int(W):
    return CubicSpline(W.t, W.data).antiderivative()(W.t)


def data_iint(W):
    return CubicSpline(W.t, W.data).antiderivative(2)(W.t)


def interpolate_data(W, tprime):
    """waveform_base.py:949-967 (data part)"""
    return CubicSpline(W.t, W.data)(tprime)


def norm(W):
    """waveform_base.py:19-35,535-551: sum |a|^2 per time"""
    return np.sum(W.data.real**2 + W.data.imag**2, axis=-1)


# ----------------------------------------------------------------------------- mode_calculations.py
_ladder = numba.njit(lambda ell, m: math.sqrt((ell - m) * (ell + m + 1)))


@jit
def _LdtVector(data, datadot, lm, Ldt):
    # mode_calculations.py:14-43
    for i_mode in range(lm.shape[0]):
        L = lm[i_mode, 0]
        M = lm[i_mode, 1]
        for i_time in range(data.shape[0]):
            Lp = np.conjugate(data[i_time, i_mode + 1]) * datadot[i_time, i_mode] * _ladder(L, M) if M + 1 <= L else 0.0 + 0.0j
            Lm = np.conjugate(data[i_time, i_mode - 1]) * datadot[i_time, i_mode] * _ladder(L, -M) if M - 1 >= -L else 0.0 + 0.0j
            Lz = np.conjugate(data[i_time, i_mode]) * datadot[i_time, i_mode] * M
            Ldt[i_time, 0] += 0.5 * (Lp.imag + Lm.imag)
            Ldt[i_time, 1] += -0.5 * (Lp.real - Lm.real)
            Ldt[i_time, 2] += Lz.imag


@jit
def _LVector(data1, data2, lm, Lvec):
    # mode_calculations.py:60-89
    for i_mode in range(lm.shape[0]):
        L = lm[i_mode, 0]
        M = lm[i_mode, 1]
        for i_time in range(data1.shape[0]):
            Lp = np.conjugate(data1[i_time, i_mode + 1]) * data2[i_time, i_mode] * _ladder(L, M) if M + 1 <= L else 0.0 + 0.0j
            Lm = np.conjugate(data1[i_time, i_mode - 1]) * data2[i_time, i_mode] * _ladder(L, -M) if M - 1 >= -L else 0.0 + 0.0j
            Lz = np.conjugate(data1[i_time, i_mode]) * data2[i_time, i_mode] * M
            Lvec[i_time, 0] += 0.5 * (Lp + Lm)
            Lvec[i_time, 1] += -0.5j * (Lp - Lm)
            Lvec[i_time, 2] += Lz


@jit
def _LLMatrix(data, lm, LL):
    # mode_calculations.py:209-295
    for i_mode in range(lm.shape[0]):
        L = lm[i_mode, 0]
        M = lm[i_mode, 1]
        for i_time in range(data.shape[0]):
            LpLp = np.conjugate(data[i_time, i_mode + 2]) * data[i_time, i_mode] * (_ladder(L, M + 1) * _ladder(L, M)) if M + 2 <= L else 0.0 + 0.0j
            LpLm = np.conjugate(data[i_time, i_mode]) * data[i_time, i_mode] * (_ladder(L, M - 1) * _ladder(L, -M)) if M - 1 >= -L else 0.0 + 0.0j
            LmLp = np.conjugate(data[i_time, i_mode]) * data[i_time, i_mode] * (_ladder(L, -(M + 1)) * _ladder(L, M)) if M + 1 <= L else 0.0 + 0.0j
            LmLm = np.conjugate(data[i_time, i_mode - 2]) * data[i_time, i_mode] * (_ladder(L, -(M - 1)) * _ladder(L, -M)) if M - 2 >= -L else 0.0 + 0.0j
            LpLz = np.conjugate(data[i_time, i_mode + 1]) * data[i_time, i_mode] * (_ladder(L, M) * M) if M + 1 <= L else 0.0 + 0.0j
            LzLp = np.conjugate(data[i_time, i_mode + 1]) * data[i_time, i_mode] * ((M + 1) * _ladder(L, M)) if M + 1 <= L else 0.0 + 0.0j
            LmLz = np.conjugate(data[i_time, i_mode - 1]) * data[i_time, i_mode] * (_ladder(L, -M) * M) if M - 1 >= -L else 0.0 + 0.0j
            LzLm = np.conjugate(data[i_time, i_mode - 1]) * data[i_time, i_mode] * ((M - 1) * _ladder(L, -M)) if M - 1 >= -L else 0.0 + 0.0j
            LzLz = np.conjugate(data[i_time, i_mode]) * data[i_time, i_mode] * M**2
            LxLx = 0.25 * (LpLp + LmLm + LmLp + LpLm)
            LxLy = -0.25j * (LpLp - LmLm + LmLp - LpLm)
            LxLz = 0.5 * (LpLz + LmLz)
            LyLx = -0.25j * (LpLp - LmLp + LpLm - LmLm)
            LyLy = -0.25 * (LpLp - LmLp - LpLm + LmLm)
            LyLz = -0.5j * (LpLz - LmLz)
            LzLx = 0.5 * (LzLp + LzLm)
            LzLy = -0.5j * (LzLp - LzLm)
            LL[i_time, 0, 0] += LxLx.real
            LL[i_time, 0, 1] += (LxLy + LyLx).real / 2.0
            LL[i_time, 0, 2] += (LxLz + LzLx).real / 2.0
            LL[i_time, 1, 0] += (LyLx + LxLy).real / 2.0
            LL[i_time, 1, 1] += LyLy.real
            LL[i_time, 1, 2] += (LyLz + LzLy).real / 2.0
            LL[i_time, 2, 0] += (LzLx + LxLz).real / 2.0
            LL[i_time, 2, 1] += (LzLy + LyLz).real / 2.0
            LL[i_time, 2, 2] += LzLz.real


@jit
def _LLComparisonMatrix(data1, data2, lm, LL):
    # mode_calculations.py:106-192, literally (LL[1,1] receives two terms and LL[1,2] none, as in the reference)
    for i_mode in range(lm.shape[0]):
        L = lm[i_mode, 0]
        M = lm[i_mode, 1]
        for i_time in range(data1.shape[0]):
            LpLp = np.conjugate(data1[i_time, i_mode + 2]) * data2[i_time, i_mode] * (_ladder(L, M + 1) * _ladder(L, M)) if M + 2 <= L else 0.0 + 0.0j
            LpLm = np.conjugate(data1[i_time, i_mode]) * data2[i_time, i_mode] * (_ladder(L, M - 1) * _ladder(L, -M)) if M - 1 >= -L else 0.0 + 0.0j
            LmLp = np.conjugate(data1[i_time, i_mode]) * data2[i_time, i_mode] * (_ladder(L, -(M + 1)) * _ladder(L, M)) if M + 1 <= L else 0.0 + 0.0j
            LmLm = np.conjugate(data1[i_time, i_mode - 2]) * data2[i_time, i_mode] * (_ladder(L, -(M - 1)) * _ladder(L, -M)) if M - 2 >= -L else 0.0 + 0.0j
            LpLz = np.conjugate(data1[i_time, i_mode + 1]) * data2[i_time, i_mode] * (_ladder(L, M) * M) if M + 1 <= L else 0.0 + 0.0j
            LzLp = np.conjugate(data1[i_time, i_mode + 1]) * data2[i_time, i_mode] * ((M + 1) * _ladder(L, M)) if M + 1 <= L else 0.0 + 0.0j
            LmLz = np.conjugate(data1[i_time, i_mode - 1]) * data2[i_time, i_mode] * (_ladder(L, -M) * M) if M - 1 >= -L else 0.0 + 0.0j
            LzLm = np.conjugate(data1[i_time, i_mode - 1]) * data2[i_time, i_mode] * ((M - 1) * _ladder(L, -M)) if M - 1 >= -L else 0.0 + 0.0j
            LzLz = np.conjugate(data1[i_time, i_mode]) * data2[i_time, i_mode] * M**2
            LL[i_time, 0, 0] += 0.25 * (LpLp + LmLm + LmLp + LpLm)
            LL[i_time, 0, 1] += -0.25j * (LpLp - LmLm + LmLp - LpLm)
            LL[i_time, 0, 2] += 0.5 * (LpLz + LmLz)
            LL[i_time, 1, 0] += -0.25j * (LpLp - LmLp + LpLm - LmLm)
            LL[i_time, 1, 1] += -0.25 * (LpLp - LmLp - LpLm + LmLm)
            LL[i_time, 1, 1] += -0.5j * (LpLz - LmLz)
            LL[i_time, 2, 0] += 0.5 * (LzLp + LzLm)
            LL[i_time, 2, 1] += -0.5j * (LzLp - LzLm)
            LL[i_time, 2, 2] += LzLz


def LLComparisonMatrix(W1, W2):
    # mode_calculations.py:195-206
    LL = np.zeros((W1.n_times, 3, 3), dtype=complex)
    _LLComparisonMatrix(W1.data, W2.data, W1.LM, LL)
    return LL


@jit
def _LLDominantEigenvector(dpa, dpa_i, i_index):
    # mode_calculations.py:316-363
    if (dpa_i[0] * dpa[i_index, 0] + dpa_i[1] * dpa[i_index, 1] + dpa_i[2] * dpa[i_index, 2]) < 0.0:
        dpa[i_index, 0] *= -1
        dpa[i_index, 1] *= -1
        dpa[i_index, 2] *= -1
    d = -1
    LastNorm = math.sqrt(dpa[i_index, 0] ** 2 + dpa[i_index, 1] ** 2 + dpa[i_index, 2] ** 2)
    for i in range(i_index - 1, -1, -1):
        Norm = dpa[i, 0] ** 2 + dpa[i, 1] ** 2 + dpa[i, 2] ** 2
        dNorm = (dpa[i, 0] - dpa[i - d, 0]) ** 2 + (dpa[i, 1] - dpa[i - d, 1]) ** 2 + (dpa[i, 2] - dpa[i - d, 2]) ** 2
        if dNorm > Norm:
            dpa[i, 0] *= -1
            dpa[i, 1] *= -1
            dpa[i, 2] *= -1
        if LastNorm != 0.0 and LastNorm != 1.0:
            dpa[i - d, 0] /= LastNorm
            dpa[i - d, 1] /= LastNorm
            dpa[i - d, 2] /= LastNorm
        LastNorm = math.sqrt(Norm)
    if LastNorm != 0.0 and LastNorm != 1.0:
        dpa[0, 0] /= LastNorm
        dpa[0, 1] /= LastNorm
        dpa[0, 2] /= LastNorm
    d = 1
    LastNorm = math.sqrt(dpa[i_index, 0] ** 2 + dpa[i_index, 1] ** 2 + dpa[i_index, 2] ** 2)
    for i in range(i_index + 1, dpa.shape[0]):
        Norm = dpa[i, 0] ** 2 + dpa[i, 1] ** 2 + dpa[i, 2] ** 2
        dNorm = (dpa[i, 0] - dpa[i - d, 0]) ** 2 + (dpa[i, 1] - dpa[i - d, 1]) ** 2 + (dpa[i, 2] - dpa[i - d, 2]) ** 2
        if dNorm > Norm:
            dpa[i, 0] *= -1
            dpa[i, 1] *= -1
            dpa[i, 2] *= -1
        if LastNorm != 0.0 and LastNorm != 1.0:
            dpa[i - d, 0] /= LastNorm
            dpa[i - d, 1] /= LastNorm
            dpa[i - d, 2] /= LastNorm
        LastNorm = math.sqrt(Norm)
    if LastNorm != 0.0 and LastNorm != 1.0:
        dpa[-1, 0] /= LastNorm
        dpa[-1, 1] /= LastNorm
        dpa[-1, 2] /= LastNorm


def LdtVector(W):
    Ldt = np.zeros((W.n_times, 3), dtype=float)
    _LdtVector(W.data, data_dot(W), W.LM, Ldt)
    return Ldt


def LVector(W):
    L = np.zeros((W.n_times, 3), dtype=complex)
    _LVector(W.data, W.data, W.LM, L)
    return L


def LLMatrix(W):
    LL = np.zeros((W.n_times, 3, 3), dtype=float)
    _LLMatrix(W.data, W.LM, LL)
    return LL


def LLDominantEigenvector(W, RoughDirection=np.array([0.0, 0.0, 1.0]), RoughDirectionIndex=0):
    # mode_calculations.py:366-399
    LL = LLMatrix(W)
    eigenvals, eigenvecs = np.linalg.eigh(LL)
    dpa = np.ascontiguousarray(eigenvecs[:, :, 2])
    _LLDominantEigenvector(dpa, np.asarray(RoughDirection, dtype=float), RoughDirectionIndex)
    return dpa


def angular_velocity(W):
    # mode_calculations.py:403-432 (include_frame_velocity=False)
    l = LdtVector(W)
    ll = LLMatrix(W)
    return -np.linalg.solve(ll, l[..., np.newaxis])[..., 0]


# ----------------------------------------------------------------------------- flux.py
def _to_matrix_indices(elements, ell_min):
    rows, cols, vals = zip(*((sf.LM_index(lp, mp, ell_min), sf.LM_index(l, m, ell_min), v) for lp, mp, l, m, v in elements))
    return np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64), np.array(vals)


@lru_cache(maxsize=None)
def p_z(ell_min, ell_max, s=-2):
    # flux.py:213-246
    out = []
    for ell in range(ell_min, ell_max + 1):
        for ellp in range(max(ell_min, ell - 1), min(ell_max, ell + 1) + 1):
            for m in range(-ell, ell + 1):
                if (m < -ellp) or (m > ellp):
                    continue
                cg1 = sf.clebsch_gordan(ell, m, 1, 0, ellp, m)
                cg2 = sf.clebsch_gordan(ell, -s, 1, 0, ellp, -s)
                prefac = np.sqrt((2.0 * ell + 1.0) / (2.0 * ellp + 1.0))
                out.append((ellp, m, ell, m, prefac * cg1 * cg2))
    return _to_matrix_indices(out, ell_min)


@lru_cache(maxsize=None)
def p_plusminus(ell_min, ell_max, sign, s=-2):
    # flux.py:249-294
    prefac = -1.0 * sign * np.sqrt(8.0 * np.pi / 3.0)

    def swsh_Y_mat_el(s, l3, m3, l1, m1, l2, m2):
        cg1 = sf.clebsch_gordan(l1, m1, l2, m2, l3, m3)
        cg2 = sf.clebsch_gordan(l1, 0, l2, -s, l3, -s)
        return np.sqrt((2.0 * l1 + 1.0) * (2.0 * l2 + 1.0) / (4.0 * np.pi * (2.0 * l3 + 1))) * cg1 * cg2

    out = []
    for ell in range(ell_min, ell_max + 1):
        for ellp in range(max(ell_min, ell - 1), min(ell_max, ell + 1) + 1):
            for m in range(-ell, ell + 1):
                mp = m + sign
                if (mp < -ellp) or (mp > ellp):
                    continue
                out.append((ellp, mp, ell, m, prefac * swsh_Y_mat_el(s, ellp, mp, 1, sign, ell, m)))
    return _to_matrix_indices(out, ell_min)


@lru_cache(maxsize=None)
def j_z(ell_min, ell_max):
    # flux.py:345-358
    out = [(ell, m, ell, m, 1.0j * m) for ell in range(ell_min, ell_max + 1) for m in range(-ell, ell + 1)]
    return _to_matrix_indices(out, ell_min)


@lru_cache(maxsize=None)
def j_plusminus(ell_min, ell_max, sign):
    # flux.py:361-385
    out = []
    for ell in range(ell_min, ell_max + 1):
        for m in range(-ell, ell + 1):
            mp = m + sign
            if (mp < -ell) or (mp > ell):
                continue
            out.append((ell, mp, ell, m, 1.0j * sf.ladder_operator_coefficient(ell, m * sign)))
    return _to_matrix_indices(out, ell_min)


@jit
def sparse_expectation_value(abar, rows, columns, values, b):
    # flux.py:40-78
    n_times = abar.shape[0]
    n_elements = rows.shape[0]
    expectation_value = np.zeros(n_times, dtype=numba.complex128)
    for i_time in range(n_times):
        for i_element in range(n_elements):
            expectation_value[i_time] += abar[i_time, rows[i_element]] * b[i_time, columns[i_element]] * values[i_element]
    return expectation_value


def _mev(a_data, M, b_data):
    rows, cols, vals = M
    return sparse_expectation_value(np.conj(a_data), rows, cols, vals.astype(complex), b_data)


def energy_flux(W):
    # flux.py:182-210
    hd = W.data if W.dataType == hdot else data_dot(W)
    Edot = np.einsum("ij, ij -> i", hd.conjugate(), hd).real
    Edot /= 16.0 * np.pi
    return Edot


def momentum_flux(W):
    # flux.py:303-342
    hd = W.data if W.dataType == hdot else data_dot(W)
    pdot = np.zeros((W.n_times, 3), dtype=float)
    pp = _mev(hd, p_plusminus(W.ell_min, W.ell_max, +1, -2), hd)
    pm = _mev(hd, p_plusminus(W.ell_min, W.ell_max, -1, -2), hd)
    pz = _mev(hd, p_z(W.ell_min, W.ell_max, -2), hd)
    pdot[:, 0] = 0.5 * (pp.real + pm.real)
    pdot[:, 1] = 0.5 * (pp.imag - pm.imag)
    pdot[:, 2] = pz.real
    pdot /= 16.0 * np.pi
    return pdot


def angular_momentum_flux(W, hdot_data=None):
    # flux.py:394-441
    hd = data_dot(W) if hdot_data is None else hdot_data
    jdot = np.zeros((W.n_times, 3), dtype=float)
    jp = _mev(hd, j_plusminus(W.ell_min, W.ell_max, +1), W.data)
    jm = _mev(hd, j_plusminus(W.ell_min, W.ell_max, -1), W.data)
    jz = _mev(hd, j_z(W.ell_min, W.ell_max), W.data)
    jdot[:, 0] = 0.5 * (jp.real + jm.real)
    jdot[:, 1] = 0.5 * (jp.imag - jm.imag)
    jdot[:, 2] = jz.real
    jdot /= -16.0 * np.pi
    return jdot


# ----------------------------------------------------------------------------- waveform_modes.py:478-562
def ladder_factor(operations, s, ell, eth_convention="NP"):
    op_dict = {"+": +1, "-": -1, +1: +1, -1: -1}
    convention_factor = {"NP": 1.0, "GHP": 0.5}[eth_convention]
    ladder = 1.0
    sign_factor = 1.0
    for op in reversed(operations):
        sign = op_dict[op]
        sign_factor *= sign
        ladder *= (ell - s * sign) * (ell + s * sign + 1.0) if (ell >= abs(s)) else 0.0
        ladder *= convention_factor
        s += sign
    return sign_factor * np.sqrt(ladder)


def apply_eth(W, operations, eth_convention="NP"):
    s = W.spin_weight
    mode_data = W.data.copy()
    for ell in range(W.ell_min, W.ell_max + 1):
        f = ladder_factor(operations, s, ell, eth_convention=eth_convention)
        idx = [sf.LM_index(ell, m, W.ell_min) for m in range(-ell, ell + 1)]
        mode_data[:, idx] *= f
    return mode_data


# ----------------------------------------------------------------------------- flux.py:444-747
@lru_cache(maxsize=None)
def eth_chi_z(ell_min, ell_max, s=-2):
    out = []
    for ell in range(ell_min, ell_max + 1):
        for ellp in range(max(ell_min, ell - 1), min(ell_max, ell + 1) + 1):
            cg2 = sf.clebsch_gordan(ell, -s, 1, -1, ellp, -1 - s)
            prefac = np.sqrt((2.0 * ell + 1.0) / (2.0 * ellp + 1.0))
            for m in range(-ell, ell + 1):
                if (m < -ellp) or (m > ellp):
                    continue
                cg1 = np.sqrt(2) * sf.clebsch_gordan(ell, m, 1, 0, ellp, m)
                out.append((ellp, m, ell, m, prefac * cg1 * cg2))
    return _to_matrix_indices(out, ell_min)


@lru_cache(maxsize=None)
def ethbar_chi_z(ell_min, ell_max, s=-2):
    out = []
    for ell in range(ell_min, ell_max + 1):
        for ellp in range(max(ell_min, ell - 1), min(ell_max, ell + 1) + 1):
            cg2 = sf.clebsch_gordan(ell, -s - 1, 1, 1, ellp, -s)
            prefac = np.sqrt((2.0 * ell + 1.0) / (2.0 * ellp + 1.0))
            for m in range(-ell, ell + 1):
                if (m < -ellp) or (m > ellp):
                    continue
                cg1 = np.sqrt(2) * sf.clebsch_gordan(ell, m, 1, 0, ellp, m)
                out.append((ellp, m, ell, m, prefac * cg1 * cg2))
    return _to_matrix_indices(out, ell_min)


@lru_cache(maxsize=None)
def eth_chi_plusminus(ell_min, ell_max, sign, s=-2):
    prefac = -1.0 * sign * np.sqrt(8.0 * np.pi / 3.0)

    def mat_el(s, l3, m3, l1, m1, l2, m2):
        cg1 = np.sqrt(2) * sf.clebsch_gordan(l1, m1, l2, m2, l3, m3)
        cg2 = sf.clebsch_gordan(l1, -1, l2, -s, l3, -1 - s)
        return np.sqrt((2.0 * l1 + 1.0) * (2.0 * l2 + 1.0) / (4.0 * np.pi * (2.0 * l3 + 1))) * cg1 * cg2

    out = []
    for ell in range(ell_min, ell_max + 1):
        for ellp in range(max(ell_min, ell - 1), min(ell_max, ell + 1) + 1):
            for m in range(-ell, ell + 1):
                mp = round(m + 1 * sign)
                if (mp < -ellp) or (mp > ellp):
                    continue
                out.append((ellp, mp, ell, m, prefac * mat_el(s, ellp, mp, 1, sign, ell, m)))
    return _to_matrix_indices(out, ell_min)


@lru_cache(maxsize=None)
def ethbar_chi_plusminus(ell_min, ell_max, sign, s=-2):
    prefac = -1.0 * sign * np.sqrt(8.0 * np.pi / 3.0)

    def mat_el(s, l3, m3, l1, m1, l2, m2):
        cg1 = np.sqrt(2) * sf.clebsch_gordan(l1, m1, l2, m2, l3, m3)
        cg2 = sf.clebsch_gordan(l1, 1, l2, -1 - s, l3, -s)
        return np.sqrt((2.0 * l1 + 1.0) * (2.0 * l2 + 1.0) / (4.0 * np.pi * (2.0 * l3 + 1))) * cg1 * cg2

    out = []
    for ell in range(ell_min, ell_max + 1):
        for ellp in range(max(ell_min, ell - 1), min(ell_max, ell + 1) + 1):
            for m in range(-ell, ell + 1):
                mp = round(m + 1 * sign)
                if (mp < -ellp) or (mp > ellp):
                    continue
                out.append((ellp, mp, ell, m, prefac * mat_el(s, ellp, mp, 1, sign, ell, m)))
    return _to_matrix_indices(out, ell_min)


def boost_flux(W, hdot_data=None):
    """flux.py:444-747 (Flanagan & Nichols 2016 eq. C.1), the 27 expectation values in the reference's order."""
    lo, hi = W.ell_min, W.ell_max
    Wd = replace(W, data=data_dot(W) if hdot_data is None else hdot_data, dataType=hdot)
    h_, hd_ = W.data, Wd.data
    eth_h, ethbar_h = apply_eth(W, "+"), apply_eth(W, "-")
    eth_hd, ethbar_hd = apply_eth(Wd, "+"), apply_eth(Wd, "-")
    pp = lambda s: p_plusminus(lo, hi, +1, s)
    pm = lambda s: p_plusminus(lo, hi, -1, s)
    pz = lambda s: p_z(lo, hi, s)
    out = np.zeros((W.n_times, 3), dtype=float)
    comps = {}
    for name, P, EC, EBC in (("plus", pp, eth_chi_plusminus(lo, hi, +1, -2), ethbar_chi_plusminus(lo, hi, +1, -2)),
                             ("minus", pm, eth_chi_plusminus(lo, hi, -1, -2), ethbar_chi_plusminus(lo, hi, -1, -2)),
                             ("z", pz, eth_chi_z(lo, hi, -2), ethbar_chi_z(lo, hi, -2))):
        first = (1 / 8) * (
            _mev(ethbar_hd, P(-3), ethbar_h)
            - _mev(eth_hd, P(-1), eth_h)
            + 6 * _mev(hd_, P(-2), h_)
            + _mev(ethbar_h, P(-3), ethbar_hd)
            - _mev(eth_h, P(-1), eth_hd)
            + 6 * _mev(h_, P(-2), hd_)
        )
        second = (1 / 2) * np.multiply(W.t, _mev(hd_, P(-2), hd_))
        a_, b_ = _mev(eth_hd, EC, h_), _mev(h_, EBC, eth_hd)
        if name == "z":
            comps[name] = first - second + (-1 / 4) * a_ + (1 / 4) * b_
        else:
            comps[name] = first - second - (1 / 4) * (a_ - b_)
    out[:, 0] = 0.5 * (comps["plus"] + comps["minus"]).real
    out[:, 1] = 0.5 * (comps["plus"] - comps["minus"]).imag
    out[:, 2] = comps["z"].real
    out /= -32 * np.pi
    return out
