"""Developer script: end-to-end time of configs[1] for the slab cut given in SCRIB200_SLAB_WEIGHTS (set before import)."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
w = bench.make_workload(100000); kw = bench.transformation_kwargs()
if os.environ.get("PIN_INPUT"):
    buf = torch.empty(w.data.shape, dtype=torch.complex128, pin_memory=True); buf.numpy()[...] = w.data
    w.data = buf.numpy(); w._keep = buf
for _ in range(6): out = w.transform(**kw)
gc.collect(); gc.disable(); ts = []
for _ in range(20):
    t0 = time.perf_counter(); out = w.transform(**kw); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print(f"{os.environ.get('SCRIB200_SLAB_WEIGHTS', 'default'):24s} pin={os.environ.get('PIN_INPUT','0')} median {np.median(ts):.2f} min {min(ts):.2f} mean {np.mean(ts):.2f}")
