import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
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
F = pl.synthesize(ad); up = pl.output_times(td)
ref = pl.remap_tiled(td, F, up)
for chunk in (256, 384, 512, 640, 768, 1024, 1536, 2048):
    pl.spline_chunk = chunk
    g = pl.remap_tiled(td, F, up); torch.cuda.synchronize()
    ts = []
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g = pl.remap_tiled(td, F, up); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(chunk, "%.3f ms" % min(ts), "maxdiff", float((g - ref).abs().max()))
