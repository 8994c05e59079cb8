"""Dev tool (GPU): end-to-end transform time and a check of the slabbed tail against the single-launch path."""
import os, sys, time
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
w = sb.WaveformModes(t=t, data=data, ell_min=2, ell_max=8, frameType=sb.Inertial, dataType=sb.h, r_is_scaled_out=True, m_is_scaled_out=True)
pl = P.TransformPlan(2, 8, sb.h, **kw)
td, ad = ops.to_device(t), ops.to_device(data)
u1, m1 = pl.run(td, ad)
for S in (2, 4, 8):
    u2, m2 = pl.run(td, ad, host_slabs=S)
    print("slabs", S, "max abs diff vs single launch", float(np.abs(m2 - m1.cpu().numpy()).max()), "bitwise equal", bool(np.array_equal(m2, m1.cpu().numpy())))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): pl.run(td, ad, host_slabs=S)
    torch.cuda.synchronize(); print("  run+D2H slabbed: %.2f ms" % ((time.perf_counter() - t0) * 100))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): ops.to_host(pl.run(td, ad)[1])
torch.cuda.synchronize(); print("run + to_host unslabbed: %.2f ms" % ((time.perf_counter() - t0) * 100))
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = w.transform(**kw)
    torch.cuda.synchronize(); print("  transform call %d: %.2f ms" % (it, (time.perf_counter() - t0) * 1e3))
