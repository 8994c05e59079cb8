"""Dev tool (GPU): boost_flux / poincare_fluxes timing (27 + 9 expectation values) at N x 285 and N x 77 modes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_inputs import smooth_modes
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
for lmax in (8, 16):
    t, data = smooth_modes(n_times=N, ell_max=lmax, t0=0.0, t1=0.1 * N, seed=1)
    w = sb.WaveformModes(t=t, data=data, ell_min=2, ell_max=lmax, frameType=sb.Inertial, dataType=sb.h, r_is_scaled_out=True, m_is_scaled_out=True)
    for name in ("boost_flux", "poincare_fluxes", "momentum_flux", "angular_momentum_flux"):
        f = getattr(w, name)
        f(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        print(f"ell<={lmax} N={N} {name}: {min(ts):.2f} ms (host arrays in and out)")
