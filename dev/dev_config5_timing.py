"""Dev tool (GPU): BASELINE configs[4] workflow through the public API (host arrays in/out) at a reduced length, with a
profile: fluxes + LLDominantEigenvector + angular velocity of an l <= 16 waveform."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import scri_b200 as sb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
w = sb.sample_waveforms.fake_precessing_waveform(t_0=-20.0, t_1=-20.0 + 0.1 * (N - 1), dt=0.1, ell_max=16)
print("N =", w.n_times, "modes", w.n_modes, "GB", w.data.nbytes / 1e9)
def T(name, f, n=2):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"{name}: {min(ts) * 1e3:.1f} ms", flush=True)
    return r
T("energy_flux", lambda: w.energy_flux())
T("momentum_flux", lambda: w.momentum_flux())
T("angular_momentum_flux", lambda: w.angular_momentum_flux())
T("poincare_fluxes", lambda: w.poincare_fluxes())
T("LLDominantEigenvector", lambda: w.LLDominantEigenvector())
T("angular_velocity", lambda: w.angular_velocity())
T("data_dot", lambda: w.data_dot)
T("norm", lambda: w.norm())
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); w.poincare_fluxes(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
