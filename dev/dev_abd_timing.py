"""Dev tool (GPU): AsymptoticBondiData.transform at BASELINE config 4 size (ell_max = 32), wall time from host arrays."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_inputs import real_supertranslation

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
rng = np.random.default_rng(0)
u = np.linspace(-10.0, 500.0, N)
n = (L + 1) ** 2
abd = sb.AsymptoticBondiData(u, L)
for name, s in (("psi0", 2), ("psi1", 1), ("psi2", 0), ("psi3", -1), ("psi4", -2), ("sigma", 2)):
    c = rng.normal(size=n) + 1j * rng.normal(size=n)
    w = rng.uniform(0.05, 0.5, size=n)
    d = 0.1 * c[None, :] * np.exp(1j * w[None, :] * u[:, None])
    d[:, : s * s] = 0.0
    setattr(abd, name, d)
kw = dict(supertranslation=real_supertranslation(2, seed=9), frame_rotation=[1.0, 2.0, 3.0, 4.0], boost_velocity=[0.01, 0.02, 0.03])
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = abd.transform(**kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"ABD.transform ell_max={L} N={N}: {dt * 1e3:.1f} ms  ({6 * n * N / dt:.3g} mode-timesteps/s over the six fields), output {out.psi4.ndarray.shape}", flush=True)
t0 = time.perf_counter()
m = abd.mass_aspect()
torch.cuda.synchronize(); print(f"mass_aspect (psi2 + sigma x d/dt bar sigma): {(time.perf_counter() - t0) * 1e3:.1f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); out = abd.transform(**kw); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
