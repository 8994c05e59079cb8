"""Pins the CPU oracle: the reference's own analytic tests ported onto oracle/ (the reference stores no golden
vectors and cannot be imported here), plus independent cross-checks against sympy / scipy.  CPU only."""
import math

import numpy as np
import pytest

from oracle import abd_ref as A, quat, scri_ref as R, sf, spinsfast as spf
from scri_inputs import real_supertranslation, rotor_set, smooth_modes


def constant_waveform(n_times=301):
    # scri/sample_waveforms.py:70-87: data m - i m, inertial, h
    t = np.linspace(-10.0, 100.0, n_times)
    LM = sf.LM_range(2, 8)
    data = np.zeros((t.size, LM.shape[0]), complex)
    for i, m in enumerate(LM[:, 1]):
        data[:, i] = m - 1j * m
    return R.Modes(t=t, data=data)


def test_wigner_d_against_sympy():
    from sympy.physics.wigner import wigner_d_small

    beta = 0.7
    for ell in (1, 2, 4):
        d = np.array(wigner_d_small(ell, beta).evalf(30).tolist(), dtype=float)
        mine = np.array([[sf.wigner_d_small(beta, ell, mp, m) for m in range(ell, -ell - 1, -1)] for mp in range(ell, -ell - 1, -1)], dtype=float)
        assert abs(mine - d.T).max() < 1e-14  # sf's d is the transpose of sympy's


def test_swsh_against_scipy_and_closed_form():
    from scipy.special import sph_harm_y

    th, ph = 1.1, 2.3
    Rq = quat.from_spherical_coords(th, ph)
    Y = sf.SWSH_grid(Rq, 0, 8)
    for l in range(9):
        for m in range(-l, l + 1):
            assert abs(Y[sf.LM_index(l, m, 0)] - sph_harm_y(l, m, th, ph)) < 5e-15
    # -2Y22 = sqrt(5/64pi) (1+cos th)^2 e^{2i ph}
    Y2 = sf.SWSH_grid(Rq, -2, 2)
    assert abs(Y2[sf.LM_index(2, 2, 0)] - math.sqrt(5 / (64 * math.pi)) * (1 + math.cos(th)) ** 2 * np.exp(2j * ph)) < 1e-15


def test_3j_and_cg_against_sympy():
    from sympy.physics.wigner import clebsch_gordan, wigner_3j

    for args in [(2, 2, 2, 0, 0, 0), (3, 4, 5, 1, -2, 1), (8, 4, 9, -3, 2, 1), (12, 4, 8, 5, -1, -4), (1, 8, 9, 1, -2, 1)]:
        assert abs(sf.Wigner3j(*args) - float(wigner_3j(*args))) < 1e-14
    assert abs(sf.clebsch_gordan(3, 1, 1, 0, 4, 1) - float(clebsch_gordan(3, 1, 4, 1, 0, 1))) < 1e-14


def test_wigner_D_is_a_representation():
    rng = np.random.default_rng(0)
    R1, R2 = quat.normalized(rng.normal(size=4)), quat.normalized(rng.normal(size=4))
    sp = quat.as_spinor_array
    D1, D2, D12 = (sf.Wigner_D_matrices(*sp(q), 2, 5) for q in (R1, R2, quat.mul(R1, R2)))
    off = 0
    for ell in range(2, 6):
        n = 2 * ell + 1
        A, B, C = (D[off : off + n * n].reshape(n, n) for D in (D1, D2, D12))
        assert abs(A @ B - C).max() < 2e-15
        assert abs(A @ A.conj().T - np.eye(n)).max() < 2e-15
        off += n * n


@pytest.mark.parametrize("s,lmax,nth,nph", [(-2, 8, 19, 19), (0, 4, 9, 9), (1, 5, 13, 11), (2, 8, 25, 25), (-1, 6, 14, 13)])
def test_spinsfast_round_trip(s, lmax, nth, nph):
    rng = np.random.default_rng(1)
    a = rng.normal(size=(3, (lmax + 1) ** 2)) + 1j * rng.normal(size=(3, (lmax + 1) ** 2))
    a[:, : s * s] = 0
    assert abs(spf.map2salm(spf.salm2map(a, s, lmax, nth, nph), s, lmax) - a).max() < 2e-14


def test_time_translation():
    """reference tests/test_waveform_grid.py:17-27"""
    dt = 1.469
    w1 = constant_waveform()
    w2 = R.transform(w1, time_translation=dt)
    w3 = R.transform(w1, supertranslation=[math.sqrt(4 * math.pi) * dt])
    assert np.allclose(w1.t, w2.t + dt, rtol=0.0, atol=2e-15)
    assert np.allclose(w1.data, w2.data, rtol=0.0, atol=4e-14)
    assert np.array_equal(w2.t, w3.t)
    assert np.array_equal(w2.data, w3.data)


def test_BMS_rotation():
    """reference tests/test_waveform_grid.py:30-38: transform(frame_rotation=R) == rotate_decomposition_basis(R)"""
    w1 = constant_waveform(101)
    for Rq in rotor_set():
        w2 = R.rotate_decomposition_basis(w1.copy(), Rq)
        w3 = R.transform(w1, frame_rotation=Rq)
        assert np.allclose(w2.data, w3.data, rtol=1e-15, atol=4e-13)


@pytest.mark.parametrize("s,ell,m", [(-2, 2, 2), (-2, 3, -1), (-2, 8, 5), (-2, 5, 0)])
def test_space_translation_analytic(s, ell, m):
    """reference tests/test_waveform_grid.py:41-92 (slow there; a few (ell,m) here), analytic formula
    scri/sample_waveforms.py:312-380: a beta*t mode supertranslated through 3j symbols."""
    t = np.arange(-20.0, 20.0 + 0.1, 0.1)
    n = sf.LM_total_size(2, 8)
    for trans in ([1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]):
        data = np.zeros((t.size, n), complex)
        data[:, sf.LM_index(ell, m, 2)] = t
        w1 = R.transform(R.Modes(t=t, data=data.copy(), dataType=R.psi4), space_translation=trans)
        st = np.zeros(4, complex)
        st[1:4] = -sf.vector_as_ell_1_modes(trans)
        expect = np.zeros((t.size, n), complex)
        expect[:, sf.LM_index(ell, m, 2)] = t
        for i, (lpp, mpp) in enumerate(sf.LM_range(0, 1)):
            if st[i] != 0.0:
                mp = m + mpp
                for lp in range(2, min(8, ell + lpp) + 1):
                    if lp >= abs(mp):
                        add = st[i] * math.sqrt(((2 * lpp + 1) * (2 * ell + 1) * (2 * lp + 1)) / (4 * math.pi)) * sf.Wigner3j(lpp, ell, lp, 0, -s, s) * sf.Wigner3j(lpp, ell, lp, mpp, m, -mp)
                        if (s + mp) % 2 == 1:
                            add *= -1
                        expect[:, sf.LM_index(lp, mp, 2)] += add
        i1 = np.argmin(abs(t - w1.t[0]))
        i2 = np.argmin(abs(t - w1.t[-1]))
        assert np.allclose(w1.data, expect[i1 : i2 + 1], rtol=0.0, atol=5e-14)


def test_supertranslation_inverses():
    """reference tests/test_waveform_grid.py:161-185 on smooth-in-time data (see SURVEY.md 7 on white noise)"""
    t, data = smooth_modes(n_times=401, t0=-10.0, t1=30.0, seed=4)
    data = data * 0 + (np.random.default_rng(4).normal(size=data.shape[1]) * (1 + 1j))[None, :] * t[:, None]
    w1 = R.Modes(t=t, data=data, dataType=R.psi4)
    for idx in (0, 2, 5, 12, 20):
        st = np.zeros(25, complex)
        lm = sf.LM_range(0, 4)[idx]
        if lm[1] == 0:
            st[idx] = 1.0
        else:
            st[sf.LM_index(lm[0], lm[1], 0)] = 1.0j
            st[sf.LM_index(lm[0], -lm[1], 0)] = (-1.0) ** lm[1] * -1.0j
        w2 = R.transform(R.transform(w1, supertranslation=st), supertranslation=-st)
        expect = R.interpolate_data(w1, w2.t)
        assert np.allclose(w2.data, expect, rtol=5e-10, atol=5e-12)


def test_rotation_inverse_and_invariants():
    """reference tests/test_rotations.py:14-129 (identity bit-exact, R then ~R, constant vs series)"""
    t, data = smooth_modes(n_times=50)
    W = R.Modes(t=t, data=data.copy())
    R.rotate_decomposition_basis(W, np.array([1.0, 0, 0, 0]))
    assert np.array_equal(W.data, data)
    for Rq in rotor_set(3):
        W = R.Modes(t=t, data=data.copy())
        R.rotate_decomposition_basis(W, Rq)
        Ws = R.Modes(t=t, data=data.copy())
        R.rotate_decomposition_basis(Ws, np.tile(Rq, (t.size, 1)))
        assert np.allclose(W.data, Ws.data, rtol=0, atol=1e-14)
        assert np.allclose(R.norm(W), np.sum(abs(data) ** 2, axis=1), rtol=1e-13)
        R.rotate_decomposition_basis(W, quat.conj(Rq))
        assert np.allclose(W.data, data, rtol=0, atol=5e-14)


def test_boost_flux_is_a_vector_like_the_momentum_flux():
    """The reference has no test for boost_flux (scri/flux.py:444-747): pin the restatement through covariance - under
    a constant rotation of the decomposition basis the boost flux must turn with the same orthogonal matrix as the
    momentum flux and the angular-momentum flux - and through its u-dependence: shifting the time origin by T adds
    T/4 times the momentum flux (the -(u/2)<N|chi|N> term of Flanagan & Nichols C.1 over -32 pi against <N|chi|N>/16 pi)."""
    t, data = smooth_modes(n_times=120, seed=3)
    W = R.Modes(t=t, data=data.copy())
    hd = R.data_dot(W)
    B0, p0, J0 = R.boost_flux(W, hd), R.momentum_flux(R.Modes(t=t, data=hd, dataType=R.hdot)), R.angular_momentum_flux(W, hd)
    for Rq in rotor_set(3)[-3:]:
        Wr, Wd = R.Modes(t=t, data=data.copy()), R.Modes(t=t, data=hd.copy(), dataType=R.hdot)
        R.rotate_decomposition_basis(Wr, Rq)
        R.rotate_decomposition_basis(Wd, Rq)
        p1, J1, B1 = R.momentum_flux(Wd), R.angular_momentum_flux(Wr, Wd.data), R.boost_flux(Wr, Wd.data)
        M = np.linalg.lstsq(p0, p1, rcond=None)[0]                 # p1 = p0 @ M
        assert np.allclose(M @ M.T, np.eye(3), atol=1e-12)
        assert np.allclose(J0 @ M, J1, rtol=0, atol=1e-12 * abs(J0).max())
        assert np.allclose(B0 @ M, B1, rtol=0, atol=1e-12 * abs(B0).max())
    T = 37.5
    Bs = R.boost_flux(R.Modes(t=t + T, data=data.copy()), hd)
    assert np.allclose(Bs, B0 + 0.25 * T * p0, rtol=0, atol=1e-12 * abs(Bs).max())


def test_abd_schwarzschild_transform():
    """reference tests/test_asymptoticbondidata.py:96-117: a boosted Schwarzschild ABD has four-momentum M gamma (1, -v)."""
    mass, ell_max = 1.0, 4
    u = np.linspace(0, 100, num=80)
    psi2 = np.zeros((ell_max + 1) ** 2, complex)
    psi2[0] = -mass * math.sqrt(4 * math.pi)                # kerr_schild(mass, 0, ell_max)[0]
    abd = A.ABD(u, ell_max, {"psi2": psi2})
    assert np.allclose(A.bondi_four_momentum(abd), [mass, 0, 0, 0], atol=1e-14)
    for v in [np.array([0.1, 0.0, 0.0]), np.array([0.0, 0.1, 0.0]), np.array([0.0, 0.0, 0.1])]:
        gamma = 1 / np.sqrt(1 - v @ v)
        p = A.bondi_four_momentum(A.transform(abd, boost_velocity=v))
        assert np.allclose(p, mass * gamma * np.array([1, *-v]), atol=1e-14, rtol=1e-14)
        assert np.allclose(np.sqrt(p[:, 0] ** 2 - np.sum(p[:, 1:] ** 2, axis=1)), mass, atol=1e-14)


def test_abd_shear_against_WaveformModes_strain():
    """reference tests/test_asymptoticbondidata.py:165-214: a general BMS transformation of an ABD's shear equals the
    transformation of the strain h = 2 bar(sigma) as a WaveformModes object (two independent code paths of the reference)."""
    rng = np.random.default_rng(123)
    ell_max = 4
    u = np.linspace(-10, 10, num=200)
    n = (ell_max + 1) ** 2
    c = 0.01 * (rng.uniform(size=(3, n)) - 0.5 + 1j * (rng.uniform(size=(3, n)) - 0.5))
    c[:, :4] = 0
    sigma = c[0][None, :] + 0.02 * c[1][None, :] * u[:, None] + 0.003 * c[2][None, :] * u[:, None] ** 2
    abd = A.ABD(u, ell_max, {"sigma": sigma})
    h = R.Modes(t=u, data=(2 * A.modes_bar(sigma, 2))[:, 4:], ell_min=2, ell_max=ell_max)
    alpha = real_supertranslation(ell_max, seed=5, scale=1e-2)
    Rq = quat.normalized(rng.normal(size=4))
    v = 0.01 * (rng.uniform(size=3) - 0.5)
    abdp = A.transform(abd, supertranslation=alpha, frame_rotation=Rq, boost_velocity=v)
    hp = R.transform(h, supertranslation=alpha, frame_rotation=Rq, boost_velocity=v)
    t = hp.t[(hp.t >= abdp.u[0]) & (hp.t <= abdp.u[-1])]
    from scipy.interpolate import CubicSpline

    a = CubicSpline(hp.t, hp.data)(t)
    b = CubicSpline(abdp.u, 2 * A.modes_bar(abdp.data["sigma"], 2))(t)[:, 4:]
    assert t.size > 150
    assert np.allclose(a, b, atol=4e-12, rtol=4e-12)


def test_LL_and_angular_velocity_of_rotating_mode():
    """reference tests/test_mode_calculations.py:14-126 (simple cases): a (2,2)+(2,-2) mode rotating about z
    with frequency w has dominant axis z and angular velocity (0,0,w)."""
    t = np.linspace(0.0, 20.0, 2001)
    omega = 0.3
    n = sf.LM_total_size(2, 4)
    data = np.zeros((t.size, n), complex)
    data[:, sf.LM_index(2, 2, 2)] = np.exp(-2j * omega * t)
    data[:, sf.LM_index(2, -2, 2)] = np.exp(2j * omega * t)
    W = R.Modes(t=t, data=data, ell_min=2, ell_max=4)
    dpa = R.LLDominantEigenvector(W)
    assert np.allclose(dpa, np.array([0, 0, 1.0])[None, :], atol=1e-14)
    om = R.angular_velocity(W)
    assert np.allclose(om[5:-5], np.array([0, 0, omega])[None, :], atol=1e-9)
    # covariance: rotate the decomposition basis by a constant rotor; omega rotates as a vector
    Rq = quat.normalized(np.array([1.0, 2.0, 3.0, 4.0]))
    W2 = R.rotate_decomposition_basis(R.Modes(t=t, data=data.copy(), ell_min=2, ell_max=4), Rq)
    om2 = R.angular_velocity(W2)
    expect = quat.rotate_vector(quat.conj(Rq), np.array([0, 0, omega]))
    assert np.allclose(om2[5:-5], expect[None, :], atol=1e-9)


def test_flux_against_grid_integration():
    """reference tests/test_flux.py:38-129: the CG-based sparse momentum / angular-momentum flux equals the
    explicit integral over the sphere (here with the oracle's own salm2map and a Gauss-Legendre-free
    Clenshaw-Curtis quadrature through map2salm of |hdot|^2 n^i)."""
    t, data = smooth_modes(n_times=60, ell_max=6, seed=3)
    W = R.Modes(t=t, data=data, ell_min=2, ell_max=6, dataType=R.hdot)
    pdot = R.momentum_flux(W)
    # explicit: dp^i/dt = (1/16pi) int |hdot|^2 n^i dOmega ; |hdot|^2 n^i has band limit 2*6+1
    L = 14
    nth = nph = 2 * L + 1
    a = np.zeros((t.size, (6 + 1) ** 2), complex)
    a[:, 4:] = data
    f = spf.salm2map(a, -2, 6, nth, nph)
    theta = np.linspace(0, np.pi, nth)
    phi = np.linspace(0, 2 * np.pi, nph, endpoint=False)
    nx = np.sin(theta)[:, None] * np.cos(phi)[None, :]
    ny = np.sin(theta)[:, None] * np.sin(phi)[None, :]
    nz = np.cos(theta)[:, None] * np.ones_like(phi)[None, :]
    p2 = abs(f) ** 2
    for i, nvec in enumerate((nx, ny, nz)):
        integ = spf.map2salm(p2 * nvec[None], 0, 0)[:, 0].real * math.sqrt(4 * math.pi)
        assert np.allclose(pdot[:, i], integ / (16 * math.pi), rtol=1e-12, atol=1e-13)
    # energy: (1/16pi) int |hdot|^2
    E = spf.map2salm(p2, 0, 0)[:, 0].real * math.sqrt(4 * math.pi) / (16 * math.pi)
    assert np.allclose(R.energy_flux(W), E, rtol=1e-12)


def test_golden_fixture_matches_oracle():
    """The committed golden vectors (tests/golden/make_golden.py) still come out of the oracle."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "transform_small.npz"))
    W = R.Modes(t=g["t"], data=g["data"])
    out = R.transform(W, supertranslation=g["supertranslation"], frame_rotation=g["frame_rotation"], boost_velocity=g["boost_velocity"])
    assert np.array_equal(out.t, g["out_t"])
    assert np.allclose(out.data, g["out_data"], rtol=0, atol=1e-14)


@pytest.mark.parametrize("s1,L1,s2,L2,Lo", [(2, 3, -2, 3, 4), (0, 2, 1, 3, 5), (-1, 2, -1, 2, 4), (2, 3, 2, 3, 6)])
def test_3j_product_equals_grid_product(s1, L1, s2, L2, Lo):
    """sf.Modes.multiply (3j sums; bms_charges.py:40-187) and ModesTimeSeries.grid_multiply (modes_time_series.py:142-202)
    compute the same coefficients once working_ell_max >= ell1 + ell2: two independent restatements pin each other."""
    rng = np.random.default_rng(1)

    def rnd(L, s):
        a = rng.normal(size=(2, (L + 1) ** 2)) + 1j * rng.normal(size=(2, (L + 1) ** 2))
        a[:, : s * s] = 0
        return a

    a, b = rnd(L1, s1), rnd(L2, s2)
    ref = A.grid_multiply(a, s1, b, s2, working_ell_max=L1 + L2, output_ell_max=Lo)
    out = sf.modes_multiply(a, s1, L1, b, s2, L2, Lo)
    assert np.abs(out - ref).max() < 5e-15 * np.abs(ref).max()


def test_wigner_d_high_ell_against_60_digit_sum():
    """Above l = 16 the oracle's Wigner D switches from the long-double explicit sum (which cancels: 1e-11 at l = 32)
    to the Jacobi-polynomial closed form; both branches against a 60-digit evaluation of the defining sum."""
    import mpmath as mp

    mp.mp.dps = 60

    def d_exact(l, mp_, m, beta):
        f = mp.factorial
        pre = mp.sqrt(f(l + mp_) * f(l - mp_) * f(l + m) * f(l - m))
        c, s = mp.cos(beta / 2), mp.sin(beta / 2)
        tot = mp.mpf(0)
        for k in range(max(0, m - mp_), min(l + m, l - mp_) + 1):
            tot += (-1) ** (k - m + mp_) * c ** (2 * l + m - mp_ - 2 * k) * s ** (2 * k - m + mp_) / (f(l + m - k) * f(k) * f(l - k - mp_) * f(k - m + mp_))
        return float(pre * tot)

    worst = 0.0
    for l in (8, 16, 17, 32, 64):
        for beta in (0.3, 1.1, math.pi / 2, 2.7):
            Ra, Rb = np.array(math.cos(beta / 2) + 0j), np.array(math.sin(beta / 2) + 0j)
            for mp_ in range(-l, l + 1, max(1, l // 4)):
                for m in range(-l, l + 1, max(1, l // 5)):
                    worst = max(worst, abs(sf.Wigner_D_element(Ra, Rb, l, mp_, m).real - d_exact(l, mp_, m, mp.mpf(beta))))
    assert worst < 1e-13


@pytest.mark.parametrize("bit_width", [8, 16, 32, 64])
def test_codec_oracle_pins(bit_width):
    """oracle/utilities_ref.py against the reference's own known answers (tests/test_utilities.py:20-52): the multishuffle with
    byte-wide pieces IS HDF5's byte shuffle (byte 0 of every element, byte 1 of every element, ...), every multishuffle is
    reversible, the XOR transform is reversible, and Fletcher-32 equals the literal 32-bit loop."""
    from oracle import utilities_ref as U

    dt = np.dtype(f"u{bit_width // 8}")
    rng = np.random.default_rng(1234)
    data = rng.integers(0, 2**bit_width, size=500, dtype=dt, endpoint=False)
    hdf5 = data.view(np.uint8).reshape(500, bit_width // 8).T.ravel().view(dt)
    assert np.array_equal(U.multishuffle((8,) * (bit_width // 8))(data), hdf5)
    for widths in [(1,) * bit_width, (8,) * (bit_width // 8), tuple([3, 5] + [1] * (bit_width - 8)), (bit_width,)]:
        assert np.array_equal(U.multishuffle(widths, False)(U.multishuffle(widths)(data)), data), widths
    x = rng.normal(size=(50, 6)) + 1j * rng.normal(size=(50, 6))
    assert np.array_equal(U.xor_timeseries_reverse(U.xor_timeseries(x)).view(np.uint64), x.view(np.uint64))
    assert np.array_equal(U.xor_timeseries(x)[0], x[0])
    d16 = rng.integers(0, 65536, size=1000 + bit_width, dtype=np.uint16)
    c0 = c1 = 0
    for j0 in range(0, d16.size, 360):
        for v in d16[j0 : j0 + 360]:
            c0 = (c0 + int(v)) & 0xFFFFFFFF
            c1 = (c1 + c0) & 0xFFFFFFFF
        c0 %= 65535
        c1 %= 65535
    assert int(U.fletcher32(d16)) == ((c1 << 16) | c0)
