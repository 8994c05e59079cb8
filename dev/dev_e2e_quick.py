"""Dev tool (GPU): end-to-end transform time and a check of the streaming pipeline against the single-launch path."""
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
pl = P.TransformPlan(2, 8, sb.h, **kw)
td, ad = ops.to_device(t), ops.to_device(data)
u1, m1 = pl.run(td, ad)
m1 = m1.cpu().numpy()
for n_slabs in (4, 8):
    a_d, slabs, fut = ops.to_device_slabs(data, np.complex128, n_slabs=n_slabs)
    u2, m2 = pl._run_streaming(td, a_d, slabs, t, debug_poison=True)
    fut.result()
    print("H2D slabs", n_slabs, "finite", bool(np.isfinite(m2).all()), "bitwise equal to the single-launch path", bool(np.array_equal(m2, m1)),
          "times equal", bool(torch.equal(u1, u2)))
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = w.transform(**kw)
    torch.cuda.synchronize(); print("  transform call %d: %.2f ms" % (it, (time.perf_counter() - t0) * 1e3))
print("result equal to device path:", bool(np.array_equal(out.data, m1)))

def e2e(streaming, n_slabs):
    a_d, slabs, fut = ops.to_device_slabs(data, np.complex128, n_slabs=n_slabs)
    u, m = pl.run(td, a_d, slabs=slabs, host_slabs=4, t_host=t if streaming else None)
    fut.result()
    return m
for streaming, n_slabs in ((False, 4), (True, 4), (True, 8), (True, 16), (False, 4), (True, 8)):
    e2e(streaming, n_slabs); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): e2e(streaming, n_slabs)
    torch.cuda.synchronize()
    print("H2D + run + D2H, streaming=%s, %d H2D slabs: %.2f ms" % (streaming, n_slabs, (time.perf_counter() - t0) * 100))
