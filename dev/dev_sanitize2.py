"""Dev tool (GPU): small invocations of the round-1 late kernels (K9 all shapes, K11, tile-skipping spline, slabbed tail)
for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_b200 import ops, plan as P
from scri_inputs import real_supertranslation, smooth_modes

rng = np.random.default_rng(0)
def rnd(N, L, s, lmin=0):
    a = rng.normal(size=(N, (L + 1) ** 2)) + 1j * rng.normal(size=(N, (L + 1) ** 2))
    a[:, : s * s] = 0
    return a[:, lmin * lmin:]
for shape in (0, 1, 2):
    for (s1, L1, s2, L2, Lw, Lo, N) in [(2, 3, -2, 4, 7, 3, 6), (2, 5, -2, 5, 6, 4, 9), (2, 32, -2, 32, 64, 32, 5)]:
        out = ops.modes_product(rnd(N, L1, s1), s1, 0, L1, rnd(N, L2, s2), s2, 0, L2, 2 * Lw + 1, 2 * Lw + 1, Lo, shape=shape)
        assert np.isfinite(out).all()
for (s, lmin, L, nth, nph, N) in [(-2, 0, 8, 17, 17, 5), (2, 2, 12, 25, 25, 9), (0, 0, 32, 65, 65, 6)]:
    g = ops.salm2map(rnd(N, L, s, lmin), s, L, nth, nph, ell_min=lmin, separable=True)
    b = ops.map2salm(g, s, L, nth, nph, ell_min=lmin, separable=True)
    assert np.isfinite(b).all()
N = 9000
t = np.linspace(0, 900.0, N)
_, data = smooth_modes(n_times=N, t0=0.0, t1=900.0)
kw = dict(supertranslation=real_supertranslation(4), frame_rotation=[1.0, 2.0, 3.0, 4.0], boost_velocity=[0.01, 0.02, 0.03])
w = sb.WaveformModes(t=t, data=data, ell_min=2, ell_max=8, frameType=sb.Inertial, dataType=sb.h, r_is_scaled_out=True, m_is_scaled_out=True)
out = w.transform(**kw)          # slabbed tail (n_out >= 8192), tile-skipping spline
assert np.isfinite(out.data).all()
torch.cuda.synchronize()
print("ok")
