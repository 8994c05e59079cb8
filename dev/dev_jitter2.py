"""Dev tool (GPU): locate the sporadic slow end-to-end calls: host marks + GPU event times of every call, printed for the slow ones."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
import scri_b200 as sb
from scri_b200 import ops, plan as P, waveform_grid as WG

kw = bench.transformation_kwargs()
w = bench.make_workload(100_000)
for _ in range(6): w.transform(**kw)
marks = []
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); marks.append((label, (t0 - T0) * 1e3, (time.perf_counter() - T0) * 1e3)); return r
    setattr(obj, name, g)
wrap(ops, "to_device_slabs", "to_device_slabs")
wrap(ops, "to_device", "to_device")
wrap(WG, "cached_transform_plan", "plan (cached)")
wrap(P.TransformPlan, "_run_streaming", "_run_streaming")
wrap(P.TransformPlan, "prepare", "prepare")
wrap(P.TimePrep, "resolve", "resolve")
wrap(P.TransformPlan, "_remap", "remap launch")
wrap(P.TransformPlan, "analyze_tiled", "analysis launch")
wrap(sb.WaveformModes, "__init__", "WaveformModes ctor")
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
if mode == "nogc":
    gc.collect(); gc.disable()
times = []
out = None
for i in range(120):
    marks.clear(); P.TRACE = []; ops.TIMING_EVENTS = []
    torch.cuda.synchronize(); T0 = time.perf_counter()
    out = w.transform(**kw)
    torch.cuda.synchronize(); tot = (time.perf_counter() - T0) * 1e3
    times.append(tot)
    if tot > 7.5:
        print(f"--- call {i}: {tot:.1f} ms")
        prev = 0.0
        for label, a, b in marks:
            if b - a > 0.5 or a - prev > 0.5: print(f"     {label:22s} {a:8.2f} -> {b:8.2f}   (gap before {a - prev:.2f})")
            prev = b
        ptt = None
        for label, tt in P.TRACE:
            x = (tt - T0) * 1e3
            if ptt is not None and x - ptt > 0.5: print(f"     [trace] {label:40s} {x:8.2f} (+{x - ptt:.2f})")
            ptt = x
        ev0 = ops.TIMING_EVENTS[0][1]; pe = 0.0
        for label, ev in ops.TIMING_EVENTS[1:]:
            x = ev0.elapsed_time(ev)
            if x - pe > 0.8: print(f"     [gpu] {label:44s} {x:8.2f} (+{x - pe:.2f})")
            pe = x
print(mode, "e2e ms:", " ".join(f"{x:.1f}" for x in times))
print("median %.2f mean %.2f max %.2f  slow(>7.5): %d" % (np.median(times), np.mean(times), max(times), sum(t > 7.5 for t in times)))
