"""Dev tool (GPU): BASELINE configs[0] - fake_precessing_waveform l <= 8, ~2e4 steps: to_corotating_frame, back to the
inertial frame, and a to_grid / from_grid round trip; wall times from host arrays, with a profile of the first."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import scri_b200 as sb

w = sb.sample_waveforms.fake_precessing_waveform(t_0=-20.0, t_1=2000.0, dt=0.1, ell_max=8)
print("N =", w.n_times)
ref = w.data.copy()
def T(f, n=3):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, r
ms, wc = T(lambda: w.copy().to_corotating_frame()); print(f"to_corotating_frame: {ms:.1f} ms")
ms, wi = T(lambda: wc.copy().to_inertial_frame()); print(f"to_inertial_frame:   {ms:.1f} ms, round-trip error {np.abs(wi.data - ref).max() / np.abs(ref).max():.2e}")
ms, g = T(lambda: w.to_grid()); print(f"to_grid:   {ms:.1f} ms  grid {g.n_theta}x{g.n_phi}")
ms, wb = T(lambda: sb.WaveformModes.from_grid(g, ell_max=8)); print(f"from_grid: {ms:.1f} ms, round-trip error {np.abs(wb.data - ref).max() / np.abs(ref).max():.2e}")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); w.copy().to_corotating_frame(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
