"""CPU oracle for the scri waveform-transformation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``scri_b200/`` imports this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU baseline.

What it is: a numpy/scipy/numba restatement of the reference's algorithm
(moble/scri 2024.0.13, ``/root/reference``) for the path named in BASELINE.json, and
of the un-vendored third-party arithmetic the reference calls on that path:

* ``oracle.quat``       <- numpy-quaternion   (>=2024.0.2; not in /root/reference)
* ``oracle.sf``         <- spherical_functions (>=2022.4;  not in /root/reference)
* ``oracle.spinsfast``  <- spinsfast (>=2022.4, Huffenberger & Wandelt 2010; not in /root/reference)
* ``oracle.scri_ref``   <- scri/waveform_grid.py, rotations.py, mode_calculations.py, flux.py,
                           waveform_base.py (spline calculus), sample_waveforms.py, tests/conftest.py

scipy's ``InterpolatedUnivariateSpline`` / ``CubicSpline`` are called exactly where the
reference calls them, so the spline arithmetic of the oracle *is* the reference's.

PARITY PINNING: the reference stores no golden vectors and cannot be imported in this image
(quaternion / spherical_functions / spinsfast / sxs / h5py are absent, no network), so the
oracle is pinned by porting the reference's own analytic tests (tests/test_oracle_*.py, each
citing the reference test it ports) and by independent cross-checks (sympy Wigner-d / 3j / CG,
scipy sph_harm_y).  Three behaviours stay "parity unpinned" against the real third-party code:
(1) spinsfast.map2salm on input that is not band-limited below N_theta-2 (we restate the
published H&W algorithm, theta-Nyquist weight taken once = Clenshaw-Curtis), and
(2) bit-level rounding of spherical_functions' Wigner-D (we agree to ~1e-15, not bit-for-bit), and
(3) numpy-quaternion's squad / integrate_angular_velocity (step selection and rounding; the product restates the
published algorithms and tests their defining properties - the reference's own corotating-frame bar is 1e-10).

``oracle.utilities_ref`` restates the integer stages of the RPXMB codec (scri/utilities.py:194-407) and IS pinned: by the
reference's known answer (multishuffle with byte-wide pieces == HDF5's byte shuffle) and by reversibility.
"""
