"""Developer script: host-side cost of TransformPlan.prepare and per-kernel GPU times."""
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
pl = P.TransformPlan(2, 8, sb.h, **kw)
td = ops.to_device(t); ad = ops.to_device(data)
for it in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prep = pl.prepare(td)
    e1.record()
    t1 = time.perf_counter()
    prep.resolve()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"prepare host {1e3*(t1-t0):.3f} ms, resolve wait {1e3*(t2-t1):.3f} ms, stream time {e0.elapsed_time(e1):.3f} ms")
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    up, m = pl.run(td, ad)
    torch.cuda.synchronize(); print(f"run wall {1e3*(time.perf_counter()-t0):.3f} ms")
