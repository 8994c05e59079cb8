"""Generate the golden fixtures from the CPU oracle (the reference itself cannot be imported in this image:
quaternion / spherical_functions / spinsfast are absent).  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import quat, scri_ref as R  # noqa: E402
from scri_inputs import real_supertranslation, smooth_modes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# 1. BMS transformation (supertranslation ell<=4 + rotation + boost), h data, ell 2..8, 161 time steps
t, data = smooth_modes(n_times=161, t0=0.0, t1=32.0, seed=11)
st = real_supertranslation(4)
rot = np.array([1.0, 2.0, 3.0, 4.0])
boost = np.array([0.01, 0.02, 0.03])
out = R.transform(R.Modes(t=t, data=data.copy()), supertranslation=st, frame_rotation=rot, boost_velocity=boost)
np.savez_compressed(os.path.join(HERE, "transform_small.npz"), t=t, data=data, supertranslation=st, frame_rotation=rot,
                    boost_velocity=boost, out_t=out.t, out_data=out.data)

# 2. rotation by a rotor series + mode calculations + fluxes, ell 2..6, non-uniform times
t, data = smooth_modes(n_times=120, ell_max=6, seed=12, uniform=False)
rng = np.random.default_rng(12)
Rs = quat.normalized(rng.normal(size=(t.size, 4)))
W = R.Modes(t=t, data=data.copy(), ell_min=2, ell_max=6)
rotated = R.rotate_decomposition_basis(W.copy(), Rs).data
np.savez_compressed(
    os.path.join(HERE, "modes_small.npz"), t=t, data=data, rotors=Rs, rotated=rotated, LL=R.LLMatrix(W), Ldt=R.LdtVector(W),
    Lvec=R.LVector(W), dpa=R.LLDominantEigenvector(W), omega=R.angular_velocity(W), data_dot=R.data_dot(W),
    energy_flux=R.energy_flux(W), momentum_flux=R.momentum_flux(W), angular_momentum_flux=R.angular_momentum_flux(W),
)
print("golden fixtures written to", HERE)
