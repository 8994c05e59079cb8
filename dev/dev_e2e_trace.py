"""Dev tool (GPU): time stamps of the stages inside the public w.transform(**kw) (monkeypatched timers + plan.TRACE)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_b200 import ops, plan as P, waveform_grid as WG
from scri_inputs import real_supertranslation, smooth_modes
N = 100000
t = np.linspace(0, 1e4, N)
_, data = smooth_modes(n_times=N, t0=0.0, t1=1e4)
kw = dict(supertranslation=real_supertranslation(4), frame_rotation=[1.0, 2.0, 3.0, 4.0], boost_velocity=[0.01, 0.02, 0.03])
w = sb.WaveformModes(t=t, data=data, ell_min=2, ell_max=8, frameType=sb.Inertial, dataType=sb.h, r_is_scaled_out=True, m_is_scaled_out=True)
for _ in range(5): w.transform(**kw)
marks = []
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); marks.append((label, (t0 - T0) * 1e3, (time.perf_counter() - T0) * 1e3)); return r
    setattr(obj, name, g)
wrap(ops, "to_device_slabs", "to_device_slabs")
wrap(ops, "to_device", "to_device")
wrap(ops, "to_host", "to_host")
wrap(WG, "cached_transform_plan", "plan (cached)")
wrap(P.TransformPlan, "_run_streaming", "_run_streaming")
wrap(P.TransformPlan, "prepare", "prepare")
wrap(P.TransformPlan, "_remap", "remap launch")
wrap(P.TransformPlan, "analyze_tiled", "analysis launch")
wrap(P.TimePrep, "resolve", "resolve")
wrap(sb.WaveformModes, "__init__", "WaveformModes ctor")
for trial in range(3):
    marks.clear()
    P.TRACE = []
    ops.TIMING_EVENTS = []
    torch.cuda.synchronize(); T0 = time.perf_counter()
    out = w.transform(**kw)
    torch.cuda.synchronize(); print("total %.2f ms" % ((time.perf_counter() - T0) * 1e3))
    for m in marks: print("   %-22s %7.2f -> %7.2f ms" % m)
    for label, tt in P.TRACE: print("   [trace] %-40s %7.2f ms" % (label, (tt - T0) * 1e3))
    ev0 = ops.TIMING_EVENTS[0][1]
    for label, ev in ops.TIMING_EVENTS[1:]: print("   [gpu] %-44s %7.2f ms after the first DMA was queued" % (label, ev0.elapsed_time(ev)))
print("registered host arrays:", len(ops._registered))
