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

PARITY PINNING (round 2): the UNMODIFIED reference runs in the build container.  It is pure Python + numba; only its
third-party packages are absent, and ``oracle/reference_loader.py`` puts stand-ins for exactly those on sys.path
(``oracle/refshim/``: quaternion, spherical_functions, spinsfast, h5py, sxs - built on the restatements above) before importing
``scri`` from /root/reference.  ``tests/golden/make_reference_vectors.py`` runs scri's own transform flow, numba loops, frame logic,
codec, AsymptoticBondiData and ``extrapolation.intersection`` on seeded inputs and commits inputs and outputs as
``tests/golden/reference_*.npz``; ``tests/test_reference_golden.py`` checks this oracle against them (transform bit-identical, codec bit
for bit), ``tests/test_gpu_reference_golden.py`` the CUDA path.  The reference's analytic tests ported onto the oracle
(tests/test_oracle.py) and the cross-checks against sympy (Wigner-d, 3j, CG) and scipy (sph_harm_y) pin the stand-ins themselves.

What stays "parity unpinned" (DESIGN.md section 2 has the measured sizes):
(1) the last bits of the real third-party packages - the golden vectors are scri's code running ON the stand-ins;
(2) spinsfast.map2salm on input that is not band limited (theta-Nyquist weight taken once = Clenshaw-Curtis here) - moot for
    band-limited input, up to 1e-3 on the thoroughly aliased late-time grid of BASELINE configs[1];
(3) the rotor ODE (different Dormand-Prince drivers at the same tolerance: frames are compared through a tighter solution);
(4) LLDominantEigenvector's sign rule when consecutive principal axes are more than 60 degrees apart (depends on LAPACK's sign).

``oracle.utilities_ref`` restates the integer stages of the RPXMB codec (scri/utilities.py:194-407) and is pinned bit for bit by the
reference's own numba functions run here, by its known answer (byte-wide multishuffle == HDF5's byte shuffle) and by reversibility.
"""
