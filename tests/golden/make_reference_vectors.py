"""Golden vectors from the UNMODIFIED reference (moble/scri 2024.0.13, /root/reference) run in the build container.

    python tests/golden/make_reference_vectors.py          (from the repo root; needs /root/reference)

The reference's own Python + numba code runs here (oracle/reference_loader.py); only the third-party packages it imports
(quaternion, spherical_functions, spinsfast - absent from this image and from /root/reference) are stood in for by
oracle/refshim/.  Every array below is therefore the output of scri's own transform flow / numba loops / frame logic /
codec on the seeded inputs stored beside it.  The fixtures are small and travel with the repo; tests compare the oracle
restatement (CPU) and the CUDA path (GPU) against them.  /root/reference does not exist on the GPU box.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import reference_loader  # noqa: E402

scri = reference_loader.load()
import quaternion  # noqa: E402  (the shim; importable after load())
from scri_inputs import real_supertranslation, smooth_modes  # noqa: E402

F = quaternion.as_float_array


def wm(t, data, ell_min=2, ell_max=8, dataType=None, frame=None, frameType=None):
    kw = dict(t=t.copy(), data=data.copy(), ell_min=ell_min, ell_max=ell_max, frameType=frameType or scri.Inertial,
              dataType=scri.h if dataType is None else dataType, r_is_scaled_out=True, m_is_scaled_out=True)
    if frame is not None:
        kw["frame"] = quaternion.as_quat_array(frame)
    return scri.WaveformModes(**kw)


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {sorted(arrays)}")


# ------------------------------------------------------------------------------------------------ 1. BMS transformations
def transforms():
    out = {}
    t, data = smooth_modes(n_times=161, t0=0.0, t1=32.0, seed=11)
    st = real_supertranslation(4)
    rot = np.array([1.0, 2.0, 3.0, 4.0])
    boost = np.array([0.01, 0.02, 0.03])
    out.update(t=t, data=data, supertranslation=st, frame_rotation=rot, boost_velocity=boost)
    cases = {
        "full": dict(supertranslation=st, frame_rotation=rot, boost_velocity=boost),
        "st": dict(supertranslation=st),
        "boost": dict(boost_velocity=boost),
        "rot": dict(frame_rotation=rot),
        "tt": dict(time_translation=1.3),
        "space": dict(space_translation=np.array([0.2, -0.1, 0.3])),
    }
    for name, kw in cases.items():
        for dt_name in ("h", "psi4", "sigma", "news"):
            if name not in ("full", "st") and dt_name != "h":
                continue
            dT = getattr(scri, dt_name)
            emin = 2
            w = wm(t, data, ell_min=emin, dataType=dT)
            r = w.transform(**{k: (v.copy() if hasattr(v, "copy") else v) for k, v in kw.items()})
            out[f"{name}_{dt_name}_t"] = r.t
            out[f"{name}_{dt_name}_data"] = r.data
    # the grid itself (to_grid / from_grid of the same waveform, and the transformed grid of the full case)
    w = wm(t, data)
    g = w.to_grid()
    out["grid_stride"] = 8
    out["grid_plain"] = g.data[::8]
    out["grid_plain_roundtrip"] = scri.WaveformModes.from_grid(g, ell_max=8).data
    g2 = scri.WaveformGrid.from_modes(w, **cases["full"])
    out["grid_full_t"] = g2.t
    out["grid_full"] = g2.data[::8]
    # psi3 needs psi4 ; psi2 needs psi3, psi4 (scri/waveform_grid.py:504-550)
    t2, d4 = smooth_modes(n_times=161, t0=0.0, t1=32.0, seed=21)
    _, d3 = smooth_modes(n_times=161, t0=0.0, t1=32.0, ell_min=1, seed=22)
    _, d2 = smooth_modes(n_times=161, t0=0.0, t1=32.0, ell_min=0, seed=23)
    w4 = wm(t2, d4, ell_min=2, dataType=scri.psi4)
    w3 = wm(t2, d3, ell_min=1, dataType=scri.psi3)
    w2 = wm(t2, d2, ell_min=0, dataType=scri.psi2)
    out.update(psi_t=t2, psi4_in=d4, psi3_in=d3, psi2_in=d2)
    r3 = w3.transform(psi4_modes=w4, **cases["full"])
    out["full_psi3_t"], out["full_psi3_data"] = r3.t, r3.data
    r2 = w2.transform(psi3_modes=w3, psi4_modes=w4, **cases["full"])
    out["full_psi2_t"], out["full_psi2_data"] = r2.t, r2.data
    save("reference_transform.npz", **out)


# ------------------------------------------------------------------------------------------------ 2. modes: rotation, <LL>, fluxes
def modes():
    out = {}
    t, data = smooth_modes(n_times=120, ell_max=6, seed=12, uniform=False)
    rng = np.random.default_rng(12)
    Rs = rng.normal(size=(t.size, 4))
    Rs /= np.linalg.norm(Rs, axis=1)[:, None]
    out.update(t=t, data=data, rotors=Rs)
    W = wm(t, data, ell_max=6)
    Wr = W.copy()
    Wr.rotate_decomposition_basis(quaternion.as_quat_array(Rs))
    out["rotated_series"] = Wr.data
    out["rotated_series_frame"] = F(Wr.frame)
    Wc = W.copy()
    Wc.rotate_decomposition_basis(quaternion.quaternion(*Rs[3]))
    out["rotated_constant"] = Wc.data
    Wp = W.copy()
    Wp.rotate_physical_system(quaternion.quaternion(*Rs[5]))
    out["rotated_physical"] = Wp.data
    out["LL"] = W.LLMatrix()
    out["Ldt"] = W.LdtVector()
    out["Lvec"] = W.LVector(W)
    out["LLcomparison"] = scri.LLComparisonMatrix(W, Wr)
    out["dpa"] = W.LLDominantEigenvector()
    out["dpa_rough"] = W.LLDominantEigenvector(RoughDirection=np.array([0.3, -0.2, -1.0]), RoughDirectionIndex=17)
    out["omega"] = W.angular_velocity()
    out["omega_rotated_with_frame"] = Wr.angular_velocity(include_frame_velocity=True)
    out["norm"] = W.norm()
    out["data_dot"], out["data_ddot"] = W.data_dot, W.data_ddot
    out["data_int"], out["data_iint"] = W.data_int, W.data_iint
    out["energy_flux"] = W.energy_flux()
    out["momentum_flux"] = W.momentum_flux()
    out["angular_momentum_flux"] = W.angular_momentum_flux()
    out["boost_flux"] = W.boost_flux()
    pf = W.poincare_fluxes()
    for k, v in zip(("pf_energy", "pf_momentum", "pf_angmom", "pf_boost"), pf):
        out[k] = v
    for ops in ("+", "-", "+-", "-+", "++", "--"):
        for conv in ("NP", "GHP"):
            out[f"eth_{ops}_{conv}"] = W.apply_eth(ops, eth_convention=conv).data
    tnew = np.sort(np.random.default_rng(5).uniform(t[3], t[-3], size=57))
    Wi = Wr.interpolate(tnew)
    out["interp_t"], out["interp_data"], out["interp_frame"] = tnew, Wi.data, F(Wi.frame)
    save("reference_modes.npz", **out)


# ------------------------------------------------------------------------------------------------ 3. frames
def frames():
    out = {}
    W = scri.sample_waveforms.fake_precessing_waveform(t_0=-20.0, t_1=400.0, dt=0.5, ell_max=4)
    out.update(t=W.t, data=W.data, frame_in=F(W.frame) if W.frame.size else np.empty((0, 4)))
    out["ell_min"], out["ell_max"] = W.ell_min, W.ell_max
    frame, omega = scri.corotating_frame(W, return_omega=True)
    out["corot_frame"], out["corot_omega"] = F(frame), omega
    frame_z = scri.corotating_frame(W, z_alignment_region=(0.1, 0.8))
    out["corot_frame_zaligned"] = F(frame_z)
    R0 = quaternion.quaternion(0.8, 0.2, -0.4, 0.4).normalized()
    out["R0"] = R0.components
    out["corot_frame_R0"] = F(scri.corotating_frame(W, R0=R0))
    Wc = W.copy()
    Wc, om, log_frame = Wc.to_corotating_frame(return_omega=True, truncate_log_frame=True, tolerance=1e-10)
    out["to_corot_trunc_data"], out["to_corot_trunc_log_frame"], out["to_corot_trunc_frame"] = Wc.data, log_frame, F(Wc.frame)
    Wc2 = W.copy().to_corotating_frame()
    out["to_corot_data"], out["to_corot_frame"] = Wc2.data, F(Wc2.frame)
    Wi = Wc2.copy().to_inertial_frame()
    out["to_inertial_data"] = Wi.data
    Wp = W.copy().to_coprecessing_frame()
    out["coprec_data"], out["coprec_frame"] = Wp.data, F(Wp.frame)
    Wpt = W.copy().to_coprecessing_frame(transition_times=(300.0, 360.0))
    out["coprec_tt_data"], out["coprec_tt_frame"] = Wpt.data, F(Wpt.frame)
    Wpr = W.copy().to_coprecessing_frame(RoughDirection=np.array([0.1, 0.1, -1.0]), RoughDirectionIndex=40)
    out["coprec_rough_frame"] = F(Wpr.frame)
    # third-party time-series routines on their own (shim = restatement; kept so that the product's versions are compared
    # with the same arithmetic the frame goldens above were made with)
    out["minimal_rotation"] = F(quaternion.minimal_rotation(frame, W.t, iterations=3))
    tnew = np.linspace(W.t[2], W.t[-3], 333)
    out["squad_t"], out["squad"] = tnew, F(quaternion.squad(frame, W.t, tnew))
    R_align = scri.rotations.get_alignment_of_decomposition_frame_to_modes(Wc2.copy(), 100.0)
    out["align_rotor"] = R_align.components
    save("reference_frames.npz", **out)


# ------------------------------------------------------------------------------------------------ 4. sample waveform generator
def samples():
    W = scri.sample_waveforms.fake_precessing_waveform(t_0=-20.0, t_1=200.0, dt=0.5, ell_max=8)
    save("reference_fake_precessing.npz", t=W.t, data=W.data, frame=F(W.frame) if W.frame.size else np.empty((0, 4)),
         args=np.array([-20.0, 200.0, 0.5, 8.0]))


# ------------------------------------------------------------------------------------------------ 5. codec (integer, bit exact)
def codec():
    from scri import utilities as U

    rng = np.random.default_rng(99)
    out = {}
    x = rng.normal(size=(301, 14)) * np.exp(rng.uniform(-20, 5, size=(301, 14)))
    out["x"] = x
    c = x.copy()
    out["xor"] = U.xor_timeseries(c).view(np.uint64)
    out["xor_reverse"] = U.xor_timeseries_reverse(U.xor_timeseries(x.copy())).view(np.uint64)
    c1 = x[:, 0].copy()
    out["xor_1d"] = U.xor_timeseries(c1).view(np.uint64)
    raw = rng.integers(0, 2**63, size=2000, dtype=np.uint64)
    out["raw"] = raw
    out["fletcher32_u16"] = np.array([U.fletcher32(raw.view(np.uint16)[:n].copy()) for n in (1, 2, 359, 360, 361, 8000)], dtype=np.uint64)
    widths_all = {
        "default": (8, 8, 4, 4, 4, 2) + (1,) * 34,
        "bytes": (8,) * 8,
        "bits": (1,) * 64,
        "mixed": (16, 3, 5, 7, 1, 32),
    }
    for k, widths in widths_all.items():
        sh = U.multishuffle(tuple(widths))
        un = U.multishuffle(tuple(widths), forward=False)
        out[f"shuffle_{k}"] = sh(raw.copy())
        out[f"unshuffle_{k}"] = un(raw.copy())
        out[f"widths_{k}"] = np.array(widths)
    for bits, dt in ((32, np.uint32), (16, np.uint16)):
        r = rng.integers(0, 2 ** (bits - 1), size=999).astype(dt)
        w = (bits // 4, bits // 4, bits // 8, bits // 8, bits // 4)
        out[f"raw{bits}"] = r
        out[f"shuffle{bits}"] = U.multishuffle(w)(r.copy())
        out[f"widths{bits}"] = np.array(w)
    # conjugate pairs and truncation (scri/waveform_modes.py:457-476,658-703)
    t, data = smooth_modes(n_times=64, ell_max=5, seed=31)
    W = wm(t, data, ell_max=5)
    Wp = W.copy()
    Wp.convert_to_conjugate_pairs()
    out["pairs_in"], out["pairs"] = data, Wp.data
    Wb = Wp.copy()
    Wb.convert_from_conjugate_pairs()
    out["pairs_back"] = Wb.data
    for tol in (1e-10, 1e-6):
        Wt = W.copy()
        Wt.truncate(tol=tol)
        out[f"truncate_{tol:g}"] = Wt.data
    save("reference_codec.npz", **out)


# ------------------------------------------------------------------------------------------------ 6. AsymptoticBondiData
def abd():
    out = {}
    ell_max = 4
    n = (ell_max + 1) ** 2
    rng = np.random.default_rng(77)
    u = np.linspace(-5.0, 25.0, 151)

    def rand_modes(s, scale):
        a = scale * (rng.normal(size=n) + 1j * rng.normal(size=n))
        a[: s * s] = 0.0
        return a

    sigma0, sigmadot0, sigmaddot0 = rand_modes(2, 1e-2), rand_modes(2, 1e-3), rand_modes(2, 1e-4)
    psi2 = rand_modes(0, 1e-2)
    psi2[0] = -1.0 * np.sqrt(4 * np.pi)
    psi1, psi0 = rand_modes(1, 1e-3), rand_modes(2, 1e-3)
    A = scri.AsymptoticBondiData.from_initial_values(u, ell_max, sigma0, sigmadot0, sigmaddot0, psi2, psi1, psi0)
    out.update(u=u, ell_max=ell_max, sigma0=sigma0, sigmadot0=sigmadot0, sigmaddot0=sigmaddot0, psi2_0=psi2, psi1_0=psi1, psi0_0=psi0)
    for name in ("psi0", "psi1", "psi2", "psi3", "psi4", "sigma"):
        out[f"in_{name}"] = np.array(getattr(A, name))
    st = real_supertranslation(3, seed=5, scale=5e-2)
    kw = dict(supertranslation=st, frame_rotation=np.array([1.0, -0.5, 0.25, 0.125]), boost_velocity=np.array([0.02, -0.01, 0.03]))
    out.update(supertranslation=st, frame_rotation=kw["frame_rotation"], boost_velocity=kw["boost_velocity"])
    B = A.transform(**kw)
    out["out_t"] = B.t
    for name in ("psi0", "psi1", "psi2", "psi3", "psi4", "sigma"):
        out[f"out_{name}"] = np.array(getattr(B, name))
    out["mass_aspect"] = np.array(A.mass_aspect())
    out["four_momentum"] = A.bondi_four_momentum()
    out["rest_mass"] = A.bondi_rest_mass()
    out["angular_momentum"] = A.bondi_angular_momentum()
    out["dimensionless_spin"] = A.bondi_dimensionless_spin()
    out["boost_charge"] = A.bondi_boost_charge()
    out["CoM_charge"] = A.bondi_CoM_charge()
    for kind in ("Bondi-Sachs", "Moreschi", "Geroch", "Geroch-Winicour"):
        out[f"supermomentum_{kind}"] = np.array(A.supermomentum(kind))
    out["grid_multiply"] = np.array(A.sigma.grid_multiply(A.sigma.bar.dot))
    out["multiply_max"] = np.array(A.sigma.multiply(A.sigma.bar, truncator=max))
    out["h_data"] = A.h.data
    from scri.asymptotic_bondi_data import transformations as T

    rot = quaternion.quaternion(*kw["frame_rotation"]).normalized()
    Rg = T.boosted_grid(rot, kw["boost_velocity"], 11, 11)
    out["boosted_grid"] = F(Rg)
    for nm, g in zip(("cf_k", "cf_ethk_over_k", "cf_one_over_k", "cf_one_over_k_cubed"), T.conformal_factors(kw["boost_velocity"], Rg)):
        out[nm] = np.asarray(g)
    save("reference_abd.npz", **out)


# ------------------------------------------------------------------------------------------------ 7. the codec chain of save()
def codec_chain():
    """The numerical part of scri/SpEC/file_io/corotating_paired_xor.py:save (lines 46-91 and 121-126; the HDF5 / JSON
    container around it needs h5py and is not part of the path), executed with the reference's own methods on a
    corotating-frame waveform: conjugate pairs -> truncate -> log(frame) on its lattice -> +0.0 -> XOR of successive
    instants -> Fletcher-32 of the three streams, plus the RPXMB multishuffle of the mode stream."""
    from scri.utilities import fletcher32, multishuffle, xor_timeseries

    fr = np.load(os.path.join(HERE, "reference_frames.npz"))
    tol = 1e-10
    w = wm(fr["t"], fr["to_corot_data"], ell_min=int(fr["ell_min"]), ell_max=int(fr["ell_max"]), frame=fr["to_corot_frame"],
           frameType=scri.Corotating)
    out = dict(t=fr["t"], data=fr["to_corot_data"], frame=fr["to_corot_frame"], ell_min=fr["ell_min"], ell_max=fr["ell_max"], tol=tol)
    w = w.copy()
    w.convert_to_conjugate_pairs()
    w.truncate(tol=tol)
    out["truncated"] = w.data.copy()
    log_frame = quaternion.as_float_array(np.log(w.frame))[:, 1:]
    power_of_2 = 2 ** (-np.floor(np.log2(tol / 10))).astype("int")
    log_frame = np.round(log_frame * power_of_2) / power_of_2
    w.t += 0.0
    w.data += 0.0
    log_frame += 0.0
    out["log_frame"] = log_frame.copy()
    xor_timeseries(w.t)
    xor_timeseries(w.data)
    xor_timeseries(log_frame)
    out["t_xor"] = w.t.view(np.uint64).copy()
    out["modes_xor"] = w.data.view(np.uint64).copy()
    out["log_frame_xor"] = log_frame.view(np.uint64).copy()
    out["fletcher32"] = np.array([fletcher32(out["t_xor"]), fletcher32(out["modes_xor"]), fletcher32(out["log_frame_xor"])], dtype=np.uint64)
    widths = (8, 8, 4, 4, 4, 2) + (1,) * 34
    out["widths"] = np.array(widths)
    out["modes_shuffled"] = multishuffle(tuple(widths))(out["modes_xor"].ravel().copy())
    save("reference_codec_chain.npz", **out)


# ------------------------------------------------------------------------------------------------ 8. <a|M|b> on differing layouts
def expectation():
    """scri/flux.py:81-179 with allow_LM_differ / allow_times_differ, and scri/extrapolation.py:47-125 (`intersection`)."""
    from scri.extrapolation import intersection
    from scri.flux import j_z, matrix_expectation_value, p_z

    rng = np.random.default_rng(41)
    ta, da = smooth_modes(n_times=140, ell_min=2, ell_max=6, t0=0.0, t1=60.0, seed=42, uniform=False)
    tb, db = smooth_modes(n_times=171, ell_min=3, ell_max=8, t0=5.0, t1=72.0, seed=43)
    a = wm(ta, da, ell_min=2, ell_max=6)
    b = wm(tb, db, ell_min=3, ell_max=8)
    out = dict(a_t=ta, a_data=da, b_t=tb, b_data=db)
    for name, M in (("p_z", p_z), ("j_z", j_z)):
        tt, val = matrix_expectation_value(a, M, b, allow_LM_differ=True, allow_times_differ=True)
        out[f"{name}_t"], out[f"{name}_val"] = tt, val
    b2 = wm(ta, smooth_modes(n_times=140, ell_min=3, ell_max=8, t0=0.0, t1=60.0, seed=44, uniform=False)[1], ell_min=3, ell_max=8)
    out["b2_data"] = b2.data
    out["lm_only_val"] = matrix_expectation_value(a, p_z, b2, allow_LM_differ=True)[1]
    t1 = np.sort(rng.uniform(0, 100, size=130))
    t2 = np.sort(rng.uniform(-10, 120, size=211))
    out.update(i_t1=t1, i_t2=t2, i_out=intersection(t1, t2), i_out_kw=intersection(t1, t2, min_step=0.9, min_time=5.0, max_time=70.0))
    save("reference_expectation.npz", **out)


# ------------------------------------------------------------------------------------------------ 9. coprecessing frame
def coprecessing():
    """scri/rotations.py:14-48 on a waveform sampled finely enough for the reference's sign rule to be well posed.
    mode_calculations.py:316-363 flips the raw eigenvector of step i when |v_i - v_{i-1}|^2 > |v_i|^2, i.e. when their dot
    product is below 1/2 - not below 0 - so when consecutive principal axes are more than 60 degrees apart the result depends
    on the sign LAPACK's eigh happened to give the raw vector.  In reference_frames.npz (dt = 0.5 M) that happens around the
    merger (|dot| down to 0.32); here (dt = 0.25 M) the smallest |dot| is 0.81 and the outcome is a property of the waveform."""
    out = {}
    W = scri.sample_waveforms.fake_precessing_waveform(t_0=-20.0, t_1=300.0, dt=0.25, ell_max=4)
    out.update(t=W.t, data=W.data, ell_min=W.ell_min, ell_max=W.ell_max)
    ev, evec = np.linalg.eigh(W.LLMatrix())
    out["min_abs_dot"] = np.abs(np.sum(evec[1:, :, 2] * evec[:-1, :, 2], axis=1)).min()
    out["dpa"] = W.LLDominantEigenvector(RoughDirectionIndex=W.n_times // 8)
    Wp = W.copy().to_coprecessing_frame()
    out["coprec_data"], out["coprec_frame"] = Wp.data, F(Wp.frame)
    Wpt = W.copy().to_coprecessing_frame(transition_times=(200.0, 260.0))
    out["coprec_tt_data"], out["coprec_tt_frame"] = Wpt.data, F(Wpt.frame)
    Wpr = W.copy().to_coprecessing_frame(RoughDirection=np.array([0.1, 0.1, -1.0]), RoughDirectionIndex=40)
    out["coprec_rough_frame"] = F(Wpr.frame)
    save("reference_coprecessing.npz", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["transforms", "modes", "frames", "samples", "codec", "abd", "codec_chain", "expectation", "coprecessing"]
    for name in which:
        globals()[name]()
