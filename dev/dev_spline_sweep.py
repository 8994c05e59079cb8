"""Developer script: spline_remap time at config-2 size against tile body / halo / CTA size."""
import os, sys
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
prep = pl.prepare(td); F = pl.synthesize(ad); up = prep.uprm
print("n_out", up.shape[0], "halo/body auto", prep.halo_body())
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
ref = None
for body, halo in ((112, 32), (128, 32), (144, 32), (192, 32), (224, 32), (240, 32)):
    pl.spline_body, pl.spline_halo = body, halo
    ts = []
    for it in range(4):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g = pl.remap_tiled(td, F, up, prep); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    if ref is None: ref = g.clone()
    print(f"body {body} halo {halo}: {min(ts[1:]):.3f} ms  maxdiff vs first {float((g - ref).abs().max()):.2e}")
