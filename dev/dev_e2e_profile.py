"""Dev tool (GPU): wall-clock profile of the public w.transform(**kw) at config-2 size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_inputs import real_supertranslation, smooth_modes
N = 100000
t = np.linspace(0, 1e4, N)
_, data = smooth_modes(n_times=N, t0=0.0, t1=1e4)
kw = dict(supertranslation=real_supertranslation(4), frame_rotation=[1.0, 2.0, 3.0, 4.0], boost_velocity=[0.01, 0.02, 0.03])
w = sb.WaveformModes(t=t, data=data, ell_min=2, ell_max=8, frameType=sb.Inertial, dataType=sb.h, r_is_scaled_out=True, m_is_scaled_out=True)
for _ in range(4): w.transform(**kw)
ts = []
for _ in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter(); w.transform(**kw); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("transform: min %.2f ms median %.2f ms" % (min(ts), sorted(ts)[5]))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): w.transform(**kw)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(16)

from scri_b200 import ops, plan as P
for trial in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a_d, slabs, fut = ops.to_device_slabs(data, np.complex128, n_slabs=8)
    fut.result(); slabs[-1][2].synchronize(); t1 = time.perf_counter()
    print("H2D alone: %.2f ms" % ((t1 - t0) * 1e3))
for trial in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a_d, slabs, fut = ops.to_device_slabs(data, np.complex128, n_slabs=8)
    pl = P.TransformPlan(2, 8, sb.h, r_is_scaled_out=True, **kw); t1 = time.perf_counter()
    fut.result(); slabs[-1][2].synchronize(); t2 = time.perf_counter()
    print("plan build while the H2D runs: plan %.2f ms, H2D done at %.2f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
for trial in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pl = P.TransformPlan(2, 8, sb.h, r_is_scaled_out=True, **kw); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("plan build alone: %.2f ms" % ((t1 - t0) * 1e3))
