"""CPU restatement of scri's AsymptoticBondiData.transform and ModesTimeSeries.grid_multiply.

TEST INFRASTRUCTURE (see oracle/__init__.py) - never imported by scri_b200/.

Follows /root/reference/scri/asymptotic_bondi_data/transformations.py:8-431, bms_charges.py:14-105 and
modes_time_series.py:142-202 with plain arrays: an `ABD` record holds u [N] and the six fields as [N, (ell_max+1)^2]
arrays (modes from ell = 0).  sf.Modes.evaluate / .eth / .bar and spinsfast are the restatements in oracle.sf /
oracle.spinsfast; CubicSpline is scipy's, exactly where the reference calls it.
"""
import math
from dataclasses import dataclass, field

import numpy as np
from scipy.interpolate import CubicSpline

from . import quat, sf, spinsfast
from .scri_ref import process_transformation_kwargs as _wg_kwargs

FIELDS = ("psi0", "psi1", "psi2", "psi3", "psi4", "sigma")
SPINS = {"psi0": 2, "psi1": 1, "psi2": 0, "psi3": -1, "psi4": -2, "sigma": 2}


@dataclass
class ABD:
    u: np.ndarray
    ell_max: int
    data: dict = field(default_factory=dict)

    def __post_init__(self):
        n = (self.ell_max + 1) ** 2
        for name in FIELDS:
            if name not in self.data:
                self.data[name] = np.zeros((self.u.size, n), dtype=complex)
            else:
                self.data[name] = np.broadcast_to(np.asarray(self.data[name], dtype=complex), (self.u.size, n)).copy()

    @property
    def n_times(self):
        return self.u.size


def _ells(ell_max):
    return np.concatenate([np.full(2 * ell + 1, float(ell)) for ell in range(ell_max + 1)])


def modes_eth(modes, s):
    """sf.Modes.eth (NP convention): multiply by sqrt((l-s)(l+s+1)); spin weight s -> s+1."""
    ell = _ells(int(round(math.sqrt(modes.shape[-1]))) - 1)
    return modes * np.where(ell >= abs(s), np.sqrt(np.maximum((ell - s) * (ell + s + 1), 0.0)), 0.0)


def modes_bar(modes, s):
    """sf.Modes.bar: modes of the conjugate function, bar_{l,m} = (-1)^{s+m} conj(f_{l,-m}); spin weight s -> -s."""
    L = int(round(math.sqrt(modes.shape[-1]))) - 1
    out = np.empty_like(modes)
    for ell in range(L + 1):
        for m in range(-ell, ell + 1):
            out[..., sf.LM_index(ell, m, 0)] = (-1.0) ** (s + m) * np.conj(modes[..., sf.LM_index(ell, -m, 0)])
    return out


def evaluate(modes, s, rotors):
    """sf.Modes.evaluate: sum_lm a_lm sY_lm(R) on a rotor grid [..., 4] -> [..., n_theta, n_phi] (time axes first)."""
    L = int(round(math.sqrt(modes.shape[-1]))) - 1
    Y = sf.SWSH_grid(rotors, s, L)
    return np.tensordot(modes, Y, axes=([-1], [-1]))


def _process_transformation_kwargs(input_ell_max, **kwargs):
    # transformations.py:8-97
    supertranslation = np.zeros((4,), dtype=complex)
    ell_max_supertranslation = 1
    if "supertranslation" in kwargs:
        supertranslation = np.array(kwargs.pop("supertranslation"), dtype=complex)
        if supertranslation.size <= 4:
            supertranslation = np.pad(supertranslation, (0, 4 - supertranslation.size), "constant", constant_values=(0.0,))
        ell_max_supertranslation = int(np.sqrt(len(supertranslation))) - 1
        if (ell_max_supertranslation + 1) ** 2 != len(supertranslation):
            raise ValueError("supertranslation length must be a perfect square")
        for ell in range(ell_max_supertranslation + 1):
            for m in range(ell + 1):
                i_pos = sf.LM_index(ell, m, 0)
                i_neg = sf.LM_index(ell, -m, 0)
                a = supertranslation[i_pos]
                b = supertranslation[i_neg]
                supertranslation[i_pos] = (a + (-1.0) ** m * b.conjugate()) / 2.0
                supertranslation[i_neg] = (-1.0) ** m * supertranslation[i_pos].conjugate()
    spacetime_translation = np.zeros((4,), dtype=float)
    if "spacetime_translation" in kwargs:
        st_trans = np.array(kwargs.pop("spacetime_translation"), dtype=float)
        spacetime_translation = st_trans[:]
        supertranslation[0] = sf.constant_as_ell_0_mode(spacetime_translation[0])
        supertranslation[1:4] = sf.vector_as_ell_1_modes(-spacetime_translation[1:4])
    if "space_translation" in kwargs:
        s_trans = np.array(kwargs.pop("space_translation"), dtype=float)
        spacetime_translation[1:4] = s_trans[:]
        supertranslation[1:4] = sf.vector_as_ell_1_modes(-spacetime_translation[1:4])
    if "time_translation" in kwargs:
        t_trans = kwargs.pop("time_translation")
        supertranslation[0] = sf.constant_as_ell_0_mode(t_trans)
    output_ell_max = kwargs.pop("output_ell_max", input_ell_max)
    working_ell_max = kwargs.pop("working_ell_max", 2 * input_ell_max + ell_max_supertranslation)
    if working_ell_max < input_ell_max:
        raise ValueError("working_ell_max is too small")
    frame_rotation = np.array(kwargs.pop("frame_rotation", [1, 0, 0, 0]), dtype=float)
    if quat.absq(frame_rotation) < 3e-16:
        raise ValueError("frame_rotation should be a single unit quaternion")
    frame_rotation = quat.normalized(frame_rotation)
    boost_velocity = np.array(kwargs.pop("boost_velocity", [0.0] * 3), dtype=float)
    beta = np.linalg.norm(boost_velocity)
    if boost_velocity.shape != (3,) or beta >= 1.0:
        raise ValueError("boost_velocity should be a 3-vector with magnitude strictly less than 1.0")
    return frame_rotation, boost_velocity, supertranslation, working_ell_max, output_ell_max


def boosted_grid(frame_rotation, boost_velocity, n_theta, n_phi):
    # transformations.py:100-148: the same rotor grid as waveform_grid.py:130-174 (restated in scri_ref)
    out = _wg_kwargs(0, frame_rotation=frame_rotation, boost_velocity=boost_velocity, n_theta=n_theta, n_phi=n_phi)
    return out[9]


def conformal_factors(boost_velocity, rotors):
    # transformations.py:151-196
    beta = np.linalg.norm(boost_velocity)
    gamma = 1 / math.sqrt(1 - beta**2)
    rz = quat.rotate_vector(rotors.reshape(-1, 4), np.array([0.0, 0.0, 1.0])).reshape(rotors.shape[:-1] + (3,))
    v_dot_r = np.dot(rz, boost_velocity)[np.newaxis, :, :]
    eth_v_dot_r = evaluate(np.insert(sf.vector_as_ell_1_modes(boost_velocity), 0, 0.0), 1, rotors)[np.newaxis, :, :]
    one_over_k = gamma * (1 - v_dot_r)
    k = 1.0 / one_over_k
    ethk_over_k = eth_v_dot_r / (1 - v_dot_r)
    return k, ethk_over_k, one_over_k, one_over_k**3


def transform(abd, **kwargs):
    # transformations.py:199-431
    frame_rotation, boost_velocity, supertranslation, working_ell_max, output_ell_max = _process_transformation_kwargs(abd.ell_max, **kwargs)
    n_theta = 2 * working_ell_max + 1
    n_phi = n_theta
    beta = np.linalg.norm(boost_velocity)
    gamma = 1 / math.sqrt(1 - beta**2)
    supertranslation = 0.5 * (supertranslation + modes_bar(supertranslation, 0))      # sf.Modes(...).real
    rotors = boosted_grid(frame_rotation, boost_velocity, n_theta, n_phi)
    u = abd.u
    alpha = evaluate(supertranslation, 0, rotors).real[np.newaxis, :, :]
    eth_alpha = (evaluate(modes_eth(supertranslation, 0), 1, rotors) / np.sqrt(2))[np.newaxis, :, :]
    ethe_alpha = (0.5 * evaluate(modes_eth(modes_eth(supertranslation, 0), 1), 2, rotors))[np.newaxis, :, :]
    k, ethk_over_k, one_over_k, one_over_k_cubed = conformal_factors(boost_velocity, rotors)
    z = ethk_over_k * (u[:, np.newaxis, np.newaxis] - alpha) - eth_alpha
    psi = [evaluate(abd.data[f"psi{n}"], SPINS[f"psi{n}"], rotors) for n in range(5)]
    sigma = evaluate(abd.data["sigma"], 2, rotors)
    fprime = np.empty((6, abd.n_times, n_theta, n_phi), dtype=complex)
    f = psi[4].copy(); f *= z; f += -4 * psi[3]; f *= z; f += 6 * psi[2]; f *= z; f += -4 * psi[1]; f *= z; f += psi[0]; f *= one_over_k_cubed
    fprime[0] = f
    f = -psi[4]; f *= z; f += 3 * psi[3]; f *= z; f += -3 * psi[2]; f *= z; f += psi[1]; f *= one_over_k_cubed
    fprime[1] = f
    f = psi[4].copy(); f *= z; f += -2 * psi[3]; f *= z; f += psi[2]; f *= one_over_k_cubed
    fprime[2] = f
    f = -psi[4]; f *= z; f += psi[3]; f *= one_over_k_cubed
    fprime[3] = f
    f = psi[4].copy(); f *= one_over_k_cubed
    fprime[4] = f
    f = sigma.copy(); f -= ethe_alpha; f *= one_over_k
    fprime[5] = f
    timeprime = (u - sf.constant_from_ell_0_mode(supertranslation[0]).real) / gamma
    earliest = np.max(k * (u[0] - alpha))
    latest = np.min(k * (u[-1] - alpha))
    timeprime = timeprime[(timeprime >= earliest) & (timeprime <= latest)]
    fout = np.zeros((6, timeprime.size, n_theta, n_phi), dtype=complex)
    for i in range(n_theta):
        for j in range(n_phi):
            x = k[0, i, j] * (u - alpha[0, i, j])
            fout[:, :, i, j] = CubicSpline(x, fprime[:, :, i, j], axis=1)(timeprime)
    out = ABD(timeprime, output_ell_max)
    for idx, name in enumerate(FIELDS):
        out.data[name] = spinsfast.map2salm(fout[idx], SPINS[name], output_ell_max)
    return out


def grid_multiply(a, sa, b, sb, working_ell_max=None, output_ell_max=None):
    # modes_time_series.py:142-202
    La = int(round(math.sqrt(a.shape[-1]))) - 1
    Lb = int(round(math.sqrt(b.shape[-1]))) - 1
    output_ell_max = La if output_ell_max is None else output_ell_max
    working_ell_max = La + Lb if working_ell_max is None else working_ell_max
    n = 2 * working_ell_max + 1
    ga = spinsfast.salm2map(a, sa, La, n, n)
    gb = spinsfast.salm2map(b, sb, Lb, n, n)
    prod = spinsfast.map2salm(ga * gb, sa + sb, working_ell_max)
    return prod[:, : sf.LM_index(output_ell_max, output_ell_max, 0) + 1]


def charge_vector_from_aspect(charge):
    # bms_charges.py:50-66
    four_vector = np.empty(charge.shape, dtype=float)
    four_vector[..., 0] = charge[..., 0].real
    four_vector[..., 1] = (charge[..., 1] - charge[..., 3]).real / math.sqrt(6)
    four_vector[..., 2] = (charge[..., 1] + charge[..., 3]).imag / math.sqrt(6)
    four_vector[..., 3] = charge[..., 2].real / math.sqrt(3)
    return four_vector / np.sqrt(4 * np.pi)


def bondi_four_momentum(abd, sigma_bar_dot=None):
    """bms_charges.py:14-47, 69-90: the ell < 2 part of M = -Re{psi2 + sigma d/dt(bar sigma)}."""
    mass_aspect = abd.data["psi2"].copy()
    if np.abs(abd.data["sigma"]).max() > 0:
        sbd = CubicSpline(abd.u, modes_bar(abd.data["sigma"], 2), axis=0).derivative()(abd.u) if sigma_bar_dot is None else sigma_bar_dot
        mass_aspect = mass_aspect + grid_multiply(abd.data["sigma"], 2, sbd, -2)
    mass_aspect = -0.5 * (mass_aspect + modes_bar(mass_aspect, 0))
    return charge_vector_from_aspect(mass_aspect[..., :4])


# ----------------------------------------------------------------------------- BMS charges through the 3j product
def modes_ethbar(modes, s):
    """sf.Modes.ethbar (NP convention): multiply by -sqrt((l+s)(l-s+1)); spin weight s -> s-1."""
    ell = _ells(int(round(math.sqrt(modes.shape[-1]))) - 1)
    return modes * np.where(ell >= abs(s), -np.sqrt(np.maximum((ell + s) * (ell - s + 1), 0.0)), 0.0)


def _real_part(modes):
    return 0.5 * (modes + modes_bar(modes, 0))


def bms_charges(abd):
    """bms_charges.py:77-189 with every product taken as spherical_functions' Modes.multiply does (3j sums,
    oracle.sf.modes_multiply) and truncated at ell = 1: four-momentum, angular momentum, boost charge, centre-of-mass charge."""
    L = abd.ell_max
    sig, psi1, psi2 = abd.data["sigma"], abd.data["psi1"], abd.data["psi2"]
    sbar = modes_bar(sig, 2)
    sbar_dot = CubicSpline(abd.u, sbar, axis=0).derivative()(abd.u)
    s_eth_sbar = sf.modes_multiply(sig, 2, L, modes_eth(sbar, -2) / math.sqrt(2), -1, L, 1)
    s_sbar = sf.modes_multiply(sig, 2, L, sbar, -2, L, 1)
    s_sbar_dot = sf.modes_multiply(sig, 2, L, sbar_dot, -2, L, 1)
    mass = -_real_part(psi2[:, :4] + s_sbar_dot)
    P = charge_vector_from_aspect(mass)
    J = charge_vector_from_aspect(1j * (psi1[:, :4] + s_eth_sbar))[:, 1:]
    com = -(psi1[:, :4] + s_eth_sbar + 0.5 * modes_eth(s_sbar, 0) / math.sqrt(2))
    G = charge_vector_from_aspect(com)[:, 1:]
    boost = com + abd.u[:, None] * modes_eth(_real_part(psi2[:, :4] + s_sbar_dot), 0) / math.sqrt(2)
    N = charge_vector_from_aspect(boost)[:, 1:]
    return P, J, N, G


def supermomentum(abd, kind, working_ell_max=None, integrated=False):
    """bms_charges.py:192-269."""
    sig = abd.data["sigma"]
    sbar = modes_bar(sig, 2)
    sbar_dot = CubicSpline(abd.u, sbar, axis=0).derivative()(abd.u)
    psi = abd.data["psi2"] + grid_multiply(sig, 2, sbar_dot, -2, working_ell_max=working_ell_max)
    eth2_sbar = modes_eth(modes_eth(sbar, -2), -1) / 2.0
    ethbar2_s = modes_ethbar(modes_ethbar(sig, 2), 1) / 2.0
    psi = psi + {"bs": 0.0, "m": eth2_sbar, "g": 0.5 * (eth2_sbar - ethbar2_s), "gw": -ethbar2_s}[kind]
    return -0.5 * modes_bar(psi, 0) / math.sqrt(math.pi) if integrated else psi
