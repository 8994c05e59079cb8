import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
from scri_b200 import ops
import scri_b200 as sb
from scri_b200.plan import TransformPlan
kw = bench.transformation_kwargs()
w = bench.make_workload(100_000)
def T(label, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
    print(f"{label:40s} {1e3*(time.perf_counter()-t0):8.2f} ms"); return r
for rep in range(2):
    print("--- rep", rep)
    T("to_device(data) alone", lambda: ops.to_device(w.data))
    plan = T("TransformPlan build", lambda: TransformPlan(w.ell_min, w.ell_max, w.dataType, r_is_scaled_out=True, **kw))
    a_d = T("to_device(data) after plan build", lambda: ops.to_device(w.data))
    T("to_device(data) again", lambda: ops.to_device(w.data))
    t_d = ops.to_device(w.t)
    up, m = T("plan.run", lambda: plan.run(t_d, a_d))
    T("to_host(modes)", lambda: ops.to_host(m))
    T("to_device(data) after run", lambda: ops.to_device(w.data))
    T("w.transform", lambda: w.transform(**kw))
