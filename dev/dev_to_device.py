"""Dev tool (GPU): time of ops.to_device on a 46 MB / 124 MB numpy array, call after call (staging ring on the first trip,
in-place registration on the second, DMA from then on)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scri_b200 import ops
for shape in ((100_000, 285), (100_000, 77), (1_000_000, 77)):
    a = np.ones(shape, dtype=np.complex128)
    ts = []
    for _ in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter(); d = ops.to_device(a); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(shape, f"{a.nbytes / 1e6:.0f} MB:", " ".join(f"{x:.2f}" for x in ts), "ms; registered:", len(ops._registered))
