"""Developer script: where does the end-to-end (host numpy in/out) transform time go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_b200 import ops, plan as P
from scri_inputs import real_supertranslation, smooth_modes
N = 100000
t = np.linspace(0, 1e4, N)
_, data = smooth_modes(n_times=N, t0=0.0, t1=1e4)
kw = dict(supertranslation=real_supertranslation(4), frame_rotation=[1.0, 2.0, 3.0, 4.0], boost_velocity=[0.01, 0.02, 0.03])
w = sb.WaveformModes(t=t, data=data, ell_min=2, ell_max=8, frameType=sb.Inertial, dataType=sb.h, r_is_scaled_out=True, m_is_scaled_out=True)
def T(f, n=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r
ms, pl = T(lambda: P.TransformPlan(2, 8, sb.h, **kw)); print("plan build %.2f ms" % ms)
ms, _ = T(lambda: P.process_transformation_kwargs(8, **dict(kw))); print("  kwargs %.2f ms" % ms)
ms, ad = T(lambda: ops.to_device(data)); print("H2D data (pinned staging) %.2f ms" % ms)
ms, ad = T(lambda: torch.from_numpy(data).cuda()); print("H2D data (pageable .cuda()) %.2f ms" % ms)
td = ops.to_device(t)
ms, (up, m) = T(lambda: pl.run(td, ad, t_ends=(t[0], t[-1]))); print("plan.run %.2f ms" % ms)
ms, mh = T(lambda: m.cpu().numpy()); print("D2H modes (.cpu()) %.2f ms" % ms)
pin = torch.empty(m.shape, dtype=m.dtype).pin_memory()
ms, _ = T(lambda: pin.copy_(m, non_blocking=True)); print("D2H modes into persistent pinned %.2f ms" % ms)
ms, out = T(lambda: w.transform(**kw)); print("w.transform total %.2f ms" % ms)
ms, _ = T(lambda: sb.WaveformModes(t=up.cpu().numpy(), data=mh, ell_min=2, ell_max=8, frameType=sb.Inertial, dataType=sb.h)); print("WaveformModes ctor %.2f ms" % ms)
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = w.transform(**kw)
    torch.cuda.synchronize(); print("  transform call %d: %.2f ms" % (it, (time.perf_counter() - t0) * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); out = w.transform(**kw); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
