"""The oracle against golden vectors produced by the UNMODIFIED reference run in the build container
(tests/golden/make_reference_vectors.py: scri's own transform flow / numba loops / frame logic / codec, with only the
absent third-party packages stood in for by oracle/refshim).  This is what pins oracle/scri_ref.py, oracle/frames_ref.py,
oracle/abd_ref.py and oracle/utilities_ref.py to the reference's code rather than to our reading of it.  CPU only."""
import os

import numpy as np
import pytest

from oracle import abd_ref, frames_ref as FR, quat, quat_series, scri_ref as R, utilities_ref as U

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


TRANSFORM_CASES = {
    "full": lambda g: dict(supertranslation=g["supertranslation"], frame_rotation=g["frame_rotation"], boost_velocity=g["boost_velocity"]),
    "st": lambda g: dict(supertranslation=g["supertranslation"]),
    "boost": lambda g: dict(boost_velocity=g["boost_velocity"]),
    "rot": lambda g: dict(frame_rotation=g["frame_rotation"]),
    "tt": lambda g: dict(time_translation=1.3),
    "space": lambda g: dict(space_translation=np.array([0.2, -0.1, 0.3])),
}
DATATYPES = {"h": 7, "psi4": 5, "sigma": 6, "news": 9}


@pytest.mark.parametrize("case", list(TRANSFORM_CASES))
def test_transform_matches_reference(case):
    """scri/waveform_grid.py:412-630 run by the reference itself vs oracle/scri_ref.py: same operations in the same order,
    hence agreement at rounding level (measured: bit-identical except `boost`, 4e-16)."""
    g = gold("reference_transform.npz")
    kw = TRANSFORM_CASES[case](g)
    checked = 0
    for name, dT in DATATYPES.items():
        key = f"{case}_{name}_data"
        if key not in g:
            continue
        w = R.Modes(t=g["t"], data=g["data"].copy(), dataType=dT)
        out = R.transform(w, **{k: (v.copy() if hasattr(v, "copy") else v) for k, v in kw.items()})
        assert np.array_equal(out.t, g[f"{case}_{name}_t"])
        assert rel(out.data, g[key]) < 1e-14
        checked += 1
    assert checked


def test_grid_and_weyl_mixing_match_reference():
    g = gold("reference_transform.npz")
    w = R.Modes(t=g["t"], data=g["data"].copy())
    kw = TRANSFORM_CASES["full"](g)
    grid = R.from_modes(w)
    assert rel(grid.data[:: g["grid_stride"]], g["grid_plain"]) < 1e-14
    assert rel(R.to_modes(grid, ell_max=8).data, g["grid_plain_roundtrip"]) < 1e-14
    grid2 = R.from_modes(w, **kw)
    assert np.array_equal(grid2.t, g["grid_full_t"])
    assert rel(grid2.data[:: g["grid_stride"]], g["grid_full"]) < 1e-14
    w4 = R.Modes(t=g["psi_t"], data=g["psi4_in"].copy(), ell_min=2, dataType=5)
    w3 = R.Modes(t=g["psi_t"], data=g["psi3_in"].copy(), ell_min=1, dataType=4)
    w2 = R.Modes(t=g["psi_t"], data=g["psi2_in"].copy(), ell_min=0, dataType=3)
    r3 = R.transform(w3, psi4_modes=w4, **TRANSFORM_CASES["full"](g))
    assert np.array_equal(r3.t, g["full_psi3_t"]) and rel(r3.data, g["full_psi3_data"]) < 1e-14
    r2 = R.transform(w2, psi3_modes=w3, psi4_modes=w4, **TRANSFORM_CASES["full"](g))
    assert np.array_equal(r2.t, g["full_psi2_t"]) and rel(r2.data, g["full_psi2_data"]) < 1e-14


def test_mode_loops_match_reference():
    """The reference's numba loops (rotations.py:346-392, mode_calculations.py:14-399, flux.py:40-78) run unchanged."""
    g = gold("reference_modes.npz")
    t, data, Rs = g["t"], g["data"], g["rotors"]
    W = R.Modes(t=t, data=data.copy(), ell_min=2, ell_max=6)
    assert rel(R.rotate_decomposition_basis(W.copy(), Rs).data, g["rotated_series"]) < 1e-14
    assert rel(R.rotate_decomposition_basis(W.copy(), Rs[3]).data, g["rotated_constant"]) < 1e-14
    assert rel(R.rotate_decomposition_basis(W.copy(), quat.conj(Rs[5])).data, g["rotated_physical"]) < 1e-14
    assert rel(R.LLMatrix(W), g["LL"]) < 1e-15
    assert rel(R.LdtVector(W), g["Ldt"]) < 1e-14
    assert rel(R.LVector(W), g["Lvec"]) < 1e-15
    Wr = R.rotate_decomposition_basis(W.copy(), Rs)
    assert rel(R.LLComparisonMatrix(W, Wr), g["LLcomparison"]) < 1e-14
    assert rel(R.LLDominantEigenvector(W), g["dpa"]) < 1e-13
    assert rel(R.LLDominantEigenvector(W, RoughDirection=np.array([0.3, -0.2, -1.0]), RoughDirectionIndex=17), g["dpa_rough"]) < 1e-13
    assert rel(R.angular_velocity(W), g["omega"]) < 1e-13
    assert rel(R.norm(W), g["norm"]) < 1e-15
    for name, fn in (("data_dot", R.data_dot), ("data_ddot", R.data_ddot), ("data_int", R.data_int), ("data_iint", R.data_iint)):
        assert rel(fn(W), g[name]) < 1e-14, name
    assert rel(R.energy_flux(W), g["energy_flux"]) < 1e-14
    assert rel(R.momentum_flux(W), g["momentum_flux"]) < 1e-14
    assert rel(R.angular_momentum_flux(W), g["angular_momentum_flux"]) < 1e-14
    assert rel(R.boost_flux(W), g["boost_flux"]) < 1e-13
    for ops in ("+", "-", "+-", "-+", "++", "--"):
        for conv in ("NP", "GHP"):
            assert rel(R.apply_eth(W, ops, eth_convention=conv).data, g[f"eth_{ops}_{conv}"]) < 1e-15
    d, fr = FR.interpolate(Wr, g["interp_t"], frame=g["rotated_series_frame"])
    assert rel(d, g["interp_data"]) < 1e-14 and rel(fr, g["interp_frame"]) < 1e-14


def test_frames_match_reference():
    """mode_calculations.py:435-490, rotations.py:14-103 run by the reference vs oracle/frames_ref.py.  The rotor ODE is
    integrated by the same restated routine on both sides; adaptive step selection amplifies rounding-level input
    differences to ~1e-8 over the series (measured), which is why the reference's own bar for these frames is 1e-8."""
    g = gold("reference_frames.npz")
    W = R.Modes(t=g["t"], data=g["data"].copy(), ell_min=int(g["ell_min"]), ell_max=int(g["ell_max"]))
    fr, om = FR.corotating_frame(W, return_omega=True)
    assert rel(om, g["corot_omega"]) < 1e-13 and rel(fr, g["corot_frame"]) < 1e-7
    assert rel(FR.corotating_frame(W, z_alignment_region=(0.1, 0.8)), g["corot_frame_zaligned"]) < 1e-7
    assert rel(FR.corotating_frame(W, R0=g["R0"]), g["corot_frame_R0"]) < 1e-7
    o, om, lf = FR.to_corotating_frame(W, tolerance=1e-10, truncate_log_frame=True)
    assert np.abs(lf - g["to_corot_trunc_log_frame"]).max() <= 2.0**-32          # at most one quantum of the 2^-33 lattice
    assert rel(o.data, g["to_corot_trunc_data"]) < 1e-7
    o, fr = FR.to_coprecessing_frame(W)
    assert rel(fr, g["coprec_frame"]) < 1e-12 and rel(o.data, g["coprec_data"]) < 1e-12
    o, fr = FR.to_coprecessing_frame(W, transition_times=(300.0, 360.0))
    assert rel(fr, g["coprec_tt_frame"]) < 1e-7 and rel(o.data, g["coprec_tt_data"]) < 1e-7
    fr = FR.coprecessing_frame(W, RoughDirection=np.array([0.1, 0.1, -1.0]), RoughDirectionIndex=40)
    assert rel(fr, g["coprec_rough_frame"]) < 1e-12
    assert rel(quat_series.minimal_rotation(g["corot_frame"], W.t, 3), g["minimal_rotation"]) < 1e-13
    assert rel(quat_series.squad(g["corot_frame"], W.t, g["squad_t"]), g["squad"]) < 1e-13


def test_codec_matches_reference_bit_for_bit():
    """scri/utilities.py:194-407 (numba, run unchanged) vs oracle/utilities_ref.py; scri/waveform_modes.py:457-476,658-703."""
    c = gold("reference_codec.npz")
    x = c["x"]
    assert np.array_equal(U.xor_timeseries(x.copy()).view(np.uint64), c["xor"])
    assert np.array_equal(U.xor_timeseries_reverse(U.xor_timeseries(x.copy())).view(np.uint64), c["xor_reverse"])
    assert np.array_equal(U.xor_timeseries(x[:, 0].copy()).view(np.uint64), c["xor_1d"])
    raw = c["raw"]
    got = [int(U.fletcher32(raw.view(np.uint16)[:n].copy())) for n in (1, 2, 359, 360, 361, 8000)]
    assert got == [int(v) for v in c["fletcher32_u16"]]
    for k in ("default", "bytes", "bits", "mixed"):
        w = tuple(int(v) for v in c[f"widths_{k}"])
        assert np.array_equal(U.multishuffle(w)(raw.copy()), c[f"shuffle_{k}"]), k
        assert np.array_equal(U.multishuffle(w, forward=False)(raw.copy()), c[f"unshuffle_{k}"]), k
    for bits in (32, 16):
        w = tuple(int(v) for v in c[f"widths{bits}"])
        assert np.array_equal(U.multishuffle(w)(c[f"raw{bits}"].copy()), c[f"shuffle{bits}"])
    assert np.array_equal(FR.convert_to_conjugate_pairs(c["pairs_in"], 2, 5), c["pairs"])
    assert np.array_equal(FR.convert_from_conjugate_pairs(c["pairs"], 2, 5), c["pairs_back"])
    for tol in (1e-10, 1e-6):
        assert np.array_equal(FR.truncate(c["pairs_in"], tol), c[f"truncate_{tol:g}"])


def test_abd_matches_reference():
    """scri/asymptotic_bondi_data/{transformations,bms_charges,from_initial_values}.py run by the reference vs
    oracle/abd_ref.py."""
    g = gold("reference_abd.npz")
    u = g["u"]
    fields = {k: g[f"in_{k}"] for k in ("psi0", "psi1", "psi2", "psi3", "psi4", "sigma")}
    A = abd_ref.ABD(u, int(g["ell_max"]), data=fields)
    B = abd_ref.transform(A, supertranslation=g["supertranslation"], frame_rotation=g["frame_rotation"], boost_velocity=g["boost_velocity"])
    assert np.array_equal(B.u, g["out_t"])
    for k in fields:
        assert rel(B.data[k], g[f"out_{k}"]) < 1e-12, k
    assert rel(abd_ref.bondi_four_momentum(A), g["four_momentum"]) < 1e-13
    P, J, N, G = abd_ref.bms_charges(A)
    assert rel(P, g["four_momentum"]) < 1e-13
    assert rel(J, g["angular_momentum"]) < 1e-12
    assert rel(N, g["boost_charge"]) < 1e-12
    assert rel(G, g["CoM_charge"]) < 1e-12
    for kind, key in (("bs", "Bondi-Sachs"), ("m", "Moreschi"), ("g", "Geroch"), ("gw", "Geroch-Winicour")):
        assert rel(abd_ref.supermomentum(A, kind), g[f"supermomentum_{key}"]) < 1e-12, kind
    from scipy.interpolate import CubicSpline

    sbd = CubicSpline(u, abd_ref.modes_bar(A.data["sigma"], 2), axis=0).derivative()(u)
    assert rel(abd_ref.grid_multiply(A.data["sigma"], 2, sbd, -2), g["grid_multiply"]) < 1e-12
    rot = quat.normalized(g["frame_rotation"])
    Rg = abd_ref.boosted_grid(rot, g["boost_velocity"], 11, 11)
    assert rel(Rg, g["boosted_grid"]) < 1e-14
    k, ethk, ok, ok3 = abd_ref.conformal_factors(g["boost_velocity"], Rg)
    assert rel(k, g["cf_k"][0]) < 1e-14 and rel(ethk, g["cf_ethk_over_k"][0]) < 1e-14 and rel(ok3, g["cf_one_over_k_cubed"][0]) < 1e-14


def test_codec_chain_matches_reference_bit_for_bit():
    """scri/SpEC/file_io/corotating_paired_xor.py:46-91,121-126 executed with the reference's own methods vs the oracle's
    restatements chained the same way."""
    c = gold("reference_codec_chain.npz")
    tol = float(c["tol"])
    d = FR.truncate(FR.convert_to_conjugate_pairs(c["data"], int(c["ell_min"]), int(c["ell_max"])), tol)
    assert np.array_equal(d.view(np.uint64), c["truncated"].view(np.uint64))
    lf = quat.log(c["frame"])[:, 1:]
    p2 = 2 ** (-np.floor(np.log2(tol / 10))).astype("int")
    lf = np.round(lf * p2) / p2 + 0.0
    assert np.array_equal(lf, c["log_frame"])
    streams = [U.xor_timeseries(c["t"] + 0.0), U.xor_timeseries(d + 0.0), U.xor_timeseries(lf.copy())]
    for got, key in zip(streams, ("t_xor", "modes_xor", "log_frame_xor")):
        assert np.array_equal(np.ascontiguousarray(got).view(np.uint64), c[key]), key
    assert [int(U.fletcher32(c[k])) for k in ("t_xor", "modes_xor", "log_frame_xor")] == [int(v) for v in c["fletcher32"]]
    w = tuple(int(v) for v in c["widths"])
    assert np.array_equal(U.multishuffle(w)(c["modes_xor"].ravel().copy()), c["modes_shuffled"])


def test_coprecessing_matches_reference_finely_sampled():
    g = gold("reference_coprecessing.npz")
    W = R.Modes(t=g["t"], data=g["data"].copy(), ell_min=int(g["ell_min"]), ell_max=int(g["ell_max"]))
    assert rel(R.LLDominantEigenvector(W, RoughDirectionIndex=W.n_times // 8), g["dpa"]) < 1e-12
    o, fr = FR.to_coprecessing_frame(W)
    assert rel(fr, g["coprec_frame"]) < 1e-12 and rel(o.data, g["coprec_data"]) < 1e-12
    o, fr = FR.to_coprecessing_frame(W, transition_times=(200.0, 260.0))
    assert rel(fr, g["coprec_tt_frame"]) < 1e-7 and rel(o.data, g["coprec_tt_data"]) < 1e-7
