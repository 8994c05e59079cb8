"""Dev tool (GPU): cProfile of WaveformModes.momentum_flux from host arrays at 1e5 x 285."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_inputs import smooth_modes
N, lmax = 100_000, 16
t, data = smooth_modes(n_times=N, ell_max=lmax, t0=0.0, t1=0.1 * N, seed=1)
w = sb.WaveformModes(t=t, data=data, ell_min=2, ell_max=lmax, frameType=sb.Inertial, dataType=sb.h, r_is_scaled_out=True, m_is_scaled_out=True)
for _ in range(3): w.momentum_flux()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): w.momentum_flux()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(22)
