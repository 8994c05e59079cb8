"""The CUDA path against golden vectors produced by the UNMODIFIED reference (tests/golden/make_reference_vectors.py ran
scri's own transform flow / numba loops / frame logic / codec in the build container and stored inputs and outputs).
Product (Python API -> ctypes -> C ABI -> sm_100a kernels) vs reference output; the oracle appears only where the size of an
ODE-integration error has to be measured (the frames).  Need a B200."""
import os

import numpy as np
import pytest

import scri_b200 as sb
from scri_b200 import _quaternion as Q

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-12        # BASELINE.json north_star: 1e-12 relative (FP64)


def gold(name):
    return np.load(os.path.join(GOLD, name))


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def wm(t, data, ell_min=2, ell_max=8, dataType=sb.h, **kw):
    return sb.WaveformModes(t=t.copy(), data=data.copy(), ell_min=ell_min, ell_max=ell_max, frameType=kw.pop("frameType", sb.Inertial),
                            dataType=dataType, r_is_scaled_out=True, m_is_scaled_out=True, **kw)


TRANSFORM_CASES = {
    "full": lambda g: dict(supertranslation=g["supertranslation"], frame_rotation=g["frame_rotation"], boost_velocity=g["boost_velocity"]),
    "st": lambda g: dict(supertranslation=g["supertranslation"]),
    "boost": lambda g: dict(boost_velocity=g["boost_velocity"]),
    "rot": lambda g: dict(frame_rotation=g["frame_rotation"]),
    "tt": lambda g: dict(time_translation=1.3),
    "space": lambda g: dict(space_translation=np.array([0.2, -0.1, 0.3])),
}
DATATYPES = {"h": sb.h, "psi4": sb.psi4, "sigma": sb.sigma, "news": sb.news}


@pytest.mark.parametrize("case", list(TRANSFORM_CASES))
def test_transform_matches_reference_output(case):
    """scri/waveform_modes.py:705-719 -> scri/waveform_grid.py:412-630 as run by the reference."""
    g = gold("reference_transform.npz")
    checked = 0
    for name, dT in DATATYPES.items():
        key = f"{case}_{name}_data"
        if key not in g:
            continue
        kw = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in TRANSFORM_CASES[case](g).items()}
        out = wm(g["t"], g["data"], dataType=dT).transform(**kw)
        assert np.array_equal(out.t, g[f"{case}_{name}_t"]), (case, name)
        assert rel(out.data, g[key]) < RTOL, (case, name, rel(out.data, g[key]))
        checked += 1
    assert checked


def test_grid_round_trip_and_weyl_mixing_match_reference_output():
    g = gold("reference_transform.npz")
    kw = TRANSFORM_CASES["full"](g)
    w = wm(g["t"], g["data"])
    grid = w.to_grid()
    stride = int(g["grid_stride"])
    assert rel(grid.data[::stride], g["grid_plain"]) < RTOL
    assert rel(sb.WaveformModes.from_grid(grid, ell_max=8).data, g["grid_plain_roundtrip"]) < RTOL
    grid2 = sb.WaveformGrid.from_modes(w, **kw)
    assert np.array_equal(grid2.t, g["grid_full_t"])
    assert rel(grid2.data[::stride], g["grid_full"]) < RTOL
    w4 = wm(g["psi_t"], g["psi4_in"], ell_min=2, dataType=sb.psi4)
    w3 = wm(g["psi_t"], g["psi3_in"], ell_min=1, dataType=sb.psi3)
    w2 = wm(g["psi_t"], g["psi2_in"], ell_min=0, dataType=sb.psi2)
    r3 = w3.transform(psi4_modes=w4, **TRANSFORM_CASES["full"](g))
    assert np.array_equal(r3.t, g["full_psi3_t"]) and rel(r3.data, g["full_psi3_data"]) < RTOL
    r2 = w2.transform(psi3_modes=w3, psi4_modes=w4, **TRANSFORM_CASES["full"](g))
    assert np.array_equal(r2.t, g["full_psi2_t"]) and rel(r2.data, g["full_psi2_data"]) < RTOL


def test_mode_loops_match_reference_output():
    """The reference's numba loops (rotations.py:346-392, mode_calculations.py:14-399, flux.py:40-78) and the time
    calculus (waveform_base.py:689-703, 949-967), run unchanged, vs kernels K4-K8."""
    g = gold("reference_modes.npz")
    t, data, Rs = g["t"], g["data"], g["rotors"]
    W = wm(t, data, ell_max=6)
    Wr = W.copy()
    Wr.rotate_decomposition_basis(Rs)
    assert rel(Wr.data, g["rotated_series"]) < 1e-13 and rel(Wr.frame, g["rotated_series_frame"]) < 1e-15
    Wc = W.copy()
    Wc.rotate_decomposition_basis(Rs[3])
    assert rel(Wc.data, g["rotated_constant"]) < 1e-13
    Wp = W.copy()
    Wp.rotate_physical_system(Rs[5])
    assert rel(Wp.data, g["rotated_physical"]) < 1e-13
    assert rel(W.LLMatrix(), g["LL"]) < 1e-13
    assert rel(W.LdtVector(), g["Ldt"]) < RTOL
    assert rel(W.LVector(), g["Lvec"]) < 1e-13
    assert rel(sb.LLComparisonMatrix(W, Wr), g["LLcomparison"]) < 1e-13
    assert rel(W.LLDominantEigenvector(), g["dpa"]) < RTOL
    assert rel(W.LLDominantEigenvector(RoughDirection=np.array([0.3, -0.2, -1.0]), RoughDirectionIndex=17), g["dpa_rough"]) < RTOL
    assert rel(W.angular_velocity(), g["omega"]) < 1e-11
    assert rel(Wr.angular_velocity(include_frame_velocity=True), g["omega_rotated_with_frame"]) < 1e-10
    assert rel(W.norm(), g["norm"]) < 1e-14
    # 120 samples at random times (intervals down to 1e-3 of the mean): the second derivative of the not-a-knot spline
    # is conditioned ~10x worse than the rest, for scipy's banded solver as for the factorisation used here
    for name, tol in (("data_dot", RTOL), ("data_ddot", 1e-11), ("data_int", RTOL), ("data_iint", RTOL)):
        assert rel(getattr(W, name), g[name]) < tol, (name, rel(getattr(W, name), g[name]))
    assert rel(W.energy_flux(), g["energy_flux"]) < RTOL
    assert rel(W.momentum_flux(), g["momentum_flux"]) < RTOL
    assert rel(W.angular_momentum_flux(), g["angular_momentum_flux"]) < RTOL
    assert rel(W.boost_flux(), g["boost_flux"]) < 1e-11
    for got, key in zip(W.poincare_fluxes(), ("pf_energy", "pf_momentum", "pf_angmom", "pf_boost")):
        assert rel(got, g[key]) < 1e-11, key
    for ops_ in ("+", "-", "+-", "-+", "++", "--"):
        for conv in ("NP", "GHP"):
            assert rel(W.apply_eth(ops_, eth_convention=conv), g[f"eth_{ops_}_{conv}"]) < 1e-15
    Wi = Wr.interpolate(g["interp_t"])
    assert rel(Wi.data, g["interp_data"]) < RTOL and rel(Wi.frame, g["interp_frame"]) < 1e-13


def test_frames_match_reference_output():
    """scri/mode_calculations.py:435-490, scri/rotations.py:14-103 as run by the reference.  Everything that involves no
    ODE agrees to rounding.  The corotating frame solves dR/dt = omega R / 2 to `tolerance` per step, by different
    Dormand-Prince drivers on the two sides (quaternion.integrate_angular_velocity's stand-in there, the library's native
    integrator here): over this 420 M series the global errors are a few 1e-8, so product and reference output are both
    measured against the oracle's integrator run with a far tighter tolerance - the product has to be at least as close
    to it as the reference output is, and the two may differ by no more than their own errors."""
    from oracle import frames_ref as FR, scri_ref as R

    g = gold("reference_frames.npz")
    lmin, lmax = int(g["ell_min"]), int(g["ell_max"])

    def W():
        return wm(g["t"], g["data"], ell_min=lmin, ell_max=lmax)

    def Wo():
        return R.Modes(t=g["t"].copy(), data=g["data"].copy(), ell_min=lmin, ell_max=lmax)

    def ode_close(mine, golden, tight):
        e_prod, e_gold = rel(mine, tight), rel(golden, tight)
        print(f"rotor ODE: product vs tight {e_prod:.2e}, reference output vs tight {e_gold:.2e}, product vs reference {rel(mine, golden):.2e}")
        return e_prod <= max(1e-8, e_gold) and rel(mine, golden) <= 1e-8 + 2 * (e_prod + e_gold)

    fr, om = sb.corotating_frame(W(), return_omega=True)
    assert rel(om, g["corot_omega"]) < 1e-11
    tight = FR.corotating_frame(Wo(), tolerance=1e-15)
    assert ode_close(fr, g["corot_frame"], tight)
    assert ode_close(sb.corotating_frame(W(), z_alignment_region=(0.1, 0.8)), g["corot_frame_zaligned"],
                     FR.corotating_frame(Wo(), z_alignment_region=(0.1, 0.8), tolerance=1e-15))
    assert ode_close(sb.corotating_frame(W(), R0=g["R0"]), g["corot_frame_R0"], FR.corotating_frame(Wo(), R0=g["R0"], tolerance=1e-15))
    w, om, lf = W().to_corotating_frame(return_omega=True, truncate_log_frame=True, tolerance=1e-10)
    assert lf.shape == g["to_corot_trunc_log_frame"].shape
    assert np.array_equal(lf * 2.0**33, np.round(lf * 2.0**33))            # on the 2^-33 lattice of tolerance = 1e-10
    # (the logarithm itself is ill-conditioned where the rotor passes near -1: the exponentials are compared)
    assert ode_close(Q.qexp(lf), Q.qexp(g["to_corot_trunc_log_frame"]), tight)       # tolerance 1e-10 per step there, too
    assert ode_close(w.frame, g["to_corot_trunc_frame"], tight)             # tolerance 1e-10 per step + one lattice step
    assert rel(w.data, R.rotate_decomposition_basis(Wo(), w.frame.copy()).data) < RTOL    # given the frame, the rotation is exact
    assert rel(w.data, g["to_corot_trunc_data"]) < 1e-5
    w, lf2 = W().to_corotating_frame(truncate_log_frame=True, tolerance=1e-10)          # the RPXMB writers' call
    assert np.array_equal(lf2, lf) and w.frameType == sb.Corotating
    w = W().to_corotating_frame()
    assert ode_close(w.frame, g["to_corot_frame"], tight)                  # same ODE, see above
    assert rel(w.data, R.rotate_decomposition_basis(Wo(), w.frame.copy()).data) < RTOL
    assert rel(w.data, g["to_corot_data"]) < 1e-6                           # modes amplify the 3e-8 frame difference by ~2 ell
    assert rel(w.to_inertial_frame().data, g["to_inertial_data"]) < 1e-12
    # coprecessing frame: this fixture is sampled so coarsely (0.5 M) that around the merger consecutive principal axes
    # are up to 71 degrees apart; the reference flips step i when dot(v_i, v_{i-1}) < 1/2 (mode_calculations.py:316-363), so
    # there its result depends on the sign LAPACK's eigh gave the raw vector.  Here only the axis itself is compared; the
    # well-posed case is test_coprecessing_frame_matches_reference_output.
    dpa = W().LLDominantEigenvector(RoughDirectionIndex=len(g["t"]) // 8)
    dpo = R.LLDominantEigenvector(Wo(), RoughDirectionIndex=len(g["t"]) // 8)
    assert np.abs(np.abs(np.sum(dpa * dpo, axis=1)) - 1).max() < 1e-12
    from scri_b200.mode_calculations import minimal_rotation

    assert rel(minimal_rotation(g["corot_frame"], g["t"], 3), g["minimal_rotation"]) < 1e-12
    assert rel(Q.squad(g["corot_frame"], g["t"], g["squad_t"]), g["squad"]) < 1e-12
    wc = wm(g["t"], g["to_corot_data"], ell_min=lmin, ell_max=lmax, frameType=sb.Corotating, frame=g["to_corot_frame"])
    Ra = sb.get_alignment_of_decomposition_frame_to_modes(wc, 100.0)
    assert min(np.abs(Ra - g["align_rotor"]).max(), np.abs(Ra + g["align_rotor"]).max()) < 1e-8


def test_fake_precessing_waveform_matches_reference_output():
    """scri/sample_waveforms.py:196-310: the bench's input generator is the reference's."""
    g = gold("reference_fake_precessing.npz")
    t0, t1, dt, L = g["args"]
    W = sb.sample_waveforms.fake_precessing_waveform(t_0=float(t0), t_1=float(t1), dt=float(dt), ell_max=int(L))
    assert np.array_equal(W.t, g["t"])
    assert rel(W.data, g["data"]) < 1e-9, rel(W.data, g["data"])      # one rotor ODE inside (precession), see above
    assert W.frame.shape[0] == g["frame"].shape[0]


def test_codec_matches_reference_output_bit_for_bit():
    """scri/utilities.py:194-407 (numba, run unchanged) and scri/waveform_modes.py:457-476,658-703 vs csrc/codec.cu."""
    from scri_b200 import utilities as ut

    c = gold("reference_codec.npz")
    x = c["x"]
    assert np.array_equal(ut.xor_timeseries(x.copy()).view(np.uint64), c["xor"])
    assert np.array_equal(ut.xor_timeseries_reverse(ut.xor_timeseries(x.copy())).view(np.uint64), c["xor_reverse"])
    assert np.array_equal(ut.xor_timeseries(x[:, 0].copy()).view(np.uint64), c["xor_1d"])
    raw = c["raw"]
    got = [int(ut.fletcher32(raw.view(np.uint16)[:n].copy())) for n in (1, 2, 359, 360, 361, 8000)]
    assert got == [int(v) for v in c["fletcher32_u16"]]
    for k in ("default", "bytes", "bits", "mixed"):
        w = tuple(int(v) for v in c[f"widths_{k}"])
        assert np.array_equal(ut.multishuffle(w)(raw.copy()), c[f"shuffle_{k}"]), k
        assert np.array_equal(ut.multishuffle(w, forward=False)(raw.copy()), c[f"unshuffle_{k}"]), k
    for bits in (32, 16):
        w = tuple(int(v) for v in c[f"widths{bits}"])
        assert np.array_equal(ut.multishuffle(w)(c[f"raw{bits}"].copy()), c[f"shuffle{bits}"])
    W = wm(np.arange(c["pairs_in"].shape[0], dtype=float), c["pairs_in"], ell_max=5)
    W.convert_to_conjugate_pairs()
    assert np.array_equal(W.data.view(np.uint64), c["pairs"].view(np.uint64))
    W.convert_from_conjugate_pairs()
    assert np.array_equal(W.data.view(np.uint64), c["pairs_back"].view(np.uint64))
    for tol in (1e-10, 1e-6):
        W = wm(np.arange(c["pairs_in"].shape[0], dtype=float), c["pairs_in"], ell_max=5)
        W.truncate(tol=tol)
        assert np.array_equal(W.data.view(np.uint64), c[f"truncate_{tol:g}"].view(np.uint64)), tol


def test_abd_matches_reference_output():
    """scri/asymptotic_bondi_data/{transformations,bms_charges}.py and scri/modes_time_series.py as run by the reference."""
    g = gold("reference_abd.npz")
    u, L = g["u"], int(g["ell_max"])
    A = sb.AsymptoticBondiData(u, L)
    for k in ("psi0", "psi1", "psi2", "psi3", "psi4", "sigma"):
        setattr(A, k, g[f"in_{k}"])
    B = A.transform(supertranslation=g["supertranslation"], frame_rotation=g["frame_rotation"], boost_velocity=g["boost_velocity"])
    assert np.array_equal(B.t, g["out_t"])
    for k in ("psi0", "psi1", "psi2", "psi3", "psi4", "sigma"):
        assert rel(getattr(B, k).ndarray, g[f"out_{k}"]) < 1e-11, (k, rel(getattr(B, k).ndarray, g[f"out_{k}"]))
    assert rel(A.mass_aspect().ndarray, g["mass_aspect"]) < RTOL
    assert rel(A.bondi_four_momentum(), g["four_momentum"]) < RTOL
    assert rel(A.bondi_rest_mass(), g["rest_mass"]) < RTOL
    assert rel(A.bondi_angular_momentum(), g["angular_momentum"]) < 1e-11
    assert rel(A.bondi_dimensionless_spin(), g["dimensionless_spin"]) < 1e-10
    assert rel(A.bondi_boost_charge(), g["boost_charge"]) < 1e-11
    assert rel(A.bondi_CoM_charge(), g["CoM_charge"]) < 1e-11
    for kind in ("Bondi-Sachs", "Moreschi", "Geroch", "Geroch-Winicour"):
        assert rel(A.supermomentum(kind).ndarray, g[f"supermomentum_{kind}"]) < 1e-11, kind
    assert rel(A.sigma.grid_multiply(A.sigma.bar.dot).ndarray, g["grid_multiply"]) < 1e-11
    assert rel(A.sigma.multiply(A.sigma.bar, truncator=max).ndarray, g["multiply_max"]) < 1e-11
    assert rel(A.h.data, g["h_data"]) < 1e-15
    from scri_b200.asymptotic_bondi_data import boosted_grid, conformal_factors

    rot = Q.qnormalized(np.asarray(g["frame_rotation"], dtype=float))
    Rg = boosted_grid(rot, g["boost_velocity"], 11, 11)
    assert rel(Rg, g["boosted_grid"]) < 1e-14
    for got, key in zip(conformal_factors(g["boost_velocity"], Rg), ("cf_k", "cf_ethk_over_k", "cf_one_over_k", "cf_one_over_k_cubed")):
        assert rel(np.broadcast_to(got, g[key].shape), g[key]) < 1e-14, key


def test_codec_chain_matches_reference_output_byte_for_byte():
    """The numerical part of scri/SpEC/file_io/corotating_paired_xor.py:save (conjugate pairs -> truncate -> log(frame) on its
    lattice -> +0.0 -> XOR of successive instants -> Fletcher-32), executed by the reference's own methods
    (tests/golden/make_reference_vectors.py:codec_chain), against scri_b200.utilities.corotating_paired_xor_encode: every
    stream the reference would hand to HDF5 is reproduced byte for byte, and so are the checksums and the RPXMB
    multishuffle of the mode stream."""
    from scri_b200 import utilities as ut

    c = gold("reference_codec_chain.npz")
    w = wm(c["t"], c["data"], ell_min=int(c["ell_min"]), ell_max=int(c["ell_max"]), frameType=sb.Corotating, frame=c["frame"])
    enc, streams, sums = ut.corotating_paired_xor_encode(w, L2norm_fractional_tolerance=float(c["tol"]))
    assert np.array_equal(streams["time"], c["t_xor"])
    assert np.array_equal(streams["modes"], c["modes_xor"])
    assert np.array_equal(streams["log_frame"], c["log_frame_xor"])
    assert [sums["time"], sums["modes"], sums["log_frame"]] == [int(v) for v in c["fletcher32"]]
    widths = tuple(int(v) for v in c["widths"])
    assert np.array_equal(ut.multishuffle(widths)(streams["modes"].ravel().copy()), c["modes_shuffled"])
    assert np.array_equal(w.data, c["data"])                     # the caller's waveform is untouched, as in the reference
    # and back: un-XOR, from conjugate pairs -> the truncated modes
    back = ut.xor_timeseries_reverse(streams["modes"].view(np.complex128).copy())
    assert np.array_equal(back.view(np.uint64), (c["truncated"] + 0.0).view(np.uint64))


def test_matrix_expectation_value_on_differing_layouts_matches_reference_output():
    """scri/flux.py:81-179 with allow_LM_differ / allow_times_differ as run by the reference: the ell ranges are clipped to
    their overlap and both waveforms interpolated onto `intersection(a.t, b.t)` before <a|M|b>."""
    from scri_b200.flux import j_z, matrix_expectation_value, p_z

    g = gold("reference_expectation.npz")
    a = wm(g["a_t"], g["a_data"], ell_min=2, ell_max=6)
    b = wm(g["b_t"], g["b_data"], ell_min=3, ell_max=8)
    for name, M in (("p_z", p_z), ("j_z", j_z)):
        tt, val = matrix_expectation_value(a, M, b, allow_LM_differ=True, allow_times_differ=True)
        assert np.array_equal(tt, g[f"{name}_t"])
        assert rel(val, g[f"{name}_val"]) < 1e-11, (name, rel(val, g[f"{name}_val"]))
    b2 = wm(g["a_t"], g["b2_data"], ell_min=3, ell_max=8)
    assert rel(matrix_expectation_value(a, p_z, b2, allow_LM_differ=True)[1], g["lm_only_val"]) < RTOL
    with pytest.raises(ValueError):
        matrix_expectation_value(a, p_z, b2)
    with pytest.raises(ValueError):
        matrix_expectation_value(a, p_z, b, allow_LM_differ=True)


def test_coprecessing_frame_matches_reference_output():
    """scri/rotations.py:14-48 as run by the reference on a waveform sampled at 0.25 M (consecutive principal axes never more
    than 36 degrees apart, so the reference's sign rule is well posed): LLDominantEigenvector with its sign walk, sqrt(-dpa z),
    minimal_rotation, and the `transition_times` branch that damps and re-integrates the frame's angular velocity."""
    g = gold("reference_coprecessing.npz")
    lmin, lmax = int(g["ell_min"]), int(g["ell_max"])
    assert float(g["min_abs_dot"]) > 0.8

    def W():
        return wm(g["t"], g["data"], ell_min=lmin, ell_max=lmax)

    assert rel(W().LLDominantEigenvector(RoughDirectionIndex=len(g["t"]) // 8), g["dpa"]) < 1e-11
    w = W().to_coprecessing_frame()
    assert w.frameType == sb.Coprecessing
    assert rel(w.frame, g["coprec_frame"]) < 1e-10 and rel(w.data, g["coprec_data"]) < 1e-10, (rel(w.frame, g["coprec_frame"]), rel(w.data, g["coprec_data"]))
    w = W().to_coprecessing_frame(transition_times=(200.0, 260.0))
    assert rel(w.frame, g["coprec_tt_frame"]) < 1e-7 and rel(w.data, g["coprec_tt_data"]) < 1e-6    # re-integrated tail: ODE tolerance
    w = W().to_coprecessing_frame(RoughDirection=np.array([0.1, 0.1, -1.0]), RoughDirectionIndex=40)
    assert rel(w.frame, g["coprec_rough_frame"]) < 1e-10
