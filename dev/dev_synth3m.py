"""Dev tool (GPU): the synthesis kernel alone at configs[1] size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import bench
from scri_b200 import ops
from scri_b200.plan import TransformPlan
kw = bench.transformation_kwargs()
w = bench.make_workload(100_000)
plan = TransformPlan(w.ell_min, w.ell_max, w.dataType, r_is_scaled_out=True, **kw)
a_d = ops.to_device(w.data)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(8):
    flush.fill_(i); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); F = plan.synthesize(a_d); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("stages", os.environ.get("SCRIB200_SYNTH3M_STAGES", "3"), "folded" if os.environ.get("SCRIB200_SYNTH_FOLDED") else "3m", "synthesis ms:", " ".join(f"{x:.3f}" for x in ts))
