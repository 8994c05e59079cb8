"""Developer script: host-side timeline of the end-to-end transform (same steps as WaveformGrid.transform)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
import scri_b200 as sb
from scri_b200 import ops
from scri_b200.plan import TransformPlan
kw = bench.transformation_kwargs()
w = bench.make_workload(100_000)
for it in range(6):
    torch.cuda.synchronize()
    T = [time.perf_counter()]
    a_d, slabs, fut = ops.to_device_slabs(w.data, np.complex128); T.append(time.perf_counter())
    plan = TransformPlan(w.ell_min, w.ell_max, w.dataType, r_is_scaled_out=True, out_ell_max=8, **kw); T.append(time.perf_counter())
    t_d = ops.to_device(w.t, np.float64); T.append(time.perf_counter())
    uprm, modes = plan.run(t_d, a_d, slabs=slabs); T.append(time.perf_counter())
    fut.result(); T.append(time.perf_counter())
    torch.cuda.current_stream().synchronize(); T.append(time.perf_counter())
    th = ops.to_host(uprm); T.append(time.perf_counter())
    mh = ops.to_host(modes); T.append(time.perf_counter())
    out = sb.WaveformModes(t=th, data=mh, ell_min=2, ell_max=8, frameType=sb.Inertial, dataType=sb.h, r_is_scaled_out=True, m_is_scaled_out=True); T.append(time.perf_counter())
    names = ["to_device_slabs(ret)", "plan build", "t H2D", "plan.run (host)", "copy thread done", "kernels done (sync)", "to_host(u')", "to_host(modes)", "WaveformModes ctor"]
    print("  ".join(f"{n} {1e3*(b-a):.2f}" for n, a, b in zip(names, T[:-1], T[1:])), f" | total {1e3*(T[-1]-T[0]):.2f} ms")
