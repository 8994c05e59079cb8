"""Developer script: small spline_remap / calculus calls for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_b200 import ops, plan as P
from scri_inputs import real_supertranslation, smooth_modes
kw = dict(supertranslation=real_supertranslation(4), frame_rotation=[1.0, 2.0, 3.0, 4.0], boost_velocity=[0.01, 0.02, 0.03])
t, data = smooth_modes(n_times=801, t0=0.0, t1=80.0)
pl = P.TransformPlan(2, 8, sb.h, **kw)
td, ad = ops.to_device(t), ops.to_device(data)
F = pl.synthesize(ad)
prep = pl.prepare(td)
up = prep.uprm
for body, halo in ((0, 0), (48, 32), (96, 64), (320, 32), (16, 128), (128, 128)):
    pl.spline_body, pl.spline_halo = body, halo
    print("body", body, "halo", halo, flush=True)
    g = pl.remap(td, F, up, prep)
    torch.cuda.synchronize()
    g = pl.remap_tiled(td, F, up, prep)
    torch.cuda.synchronize()
for kind, order in (("derivative", 1), ("derivative", 2), ("antiderivative", 1), ("antiderivative", 2)):
    print(kind, order, flush=True)
    ops.spline_calculus(t, data, kind, order)
    torch.cuda.synchronize()
for n in (4, 5, 17):
    print("n", n, flush=True)
    ops.spline_calculus(t[:n], data[:n], "evaluate", tprime=np.linspace(t[0], t[n - 1], 50))
    ops.spline_calculus(t[:n], data[:n], "derivative", 1)
    torch.cuda.synchronize()
print("done")
