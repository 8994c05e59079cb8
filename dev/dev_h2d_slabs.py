"""Developer script: cost of cutting one 124 MB H2D into slabs (weights as in WaveformGrid.from_modes), on one or two copy
streams, through torch copy_ or the library's scrib200_h2d, with and without a D2H of the same size the other way."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scri_b200 import ops, _lib
lib = _lib.load()
n = 124_000_000
pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True); pinned.fill_(1)
dst = torch.empty(n, dtype=torch.uint8, pin_memory=True); dst.fill_(2)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.ones(n, dtype=torch.uint8, device="cuda")
sA, sB, s2 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def bounds(weights):
    cum = np.concatenate([[0.0], np.cumsum(np.asarray(weights, float))]); return [int(round(n * c / cum[-1])) & ~15 for c in cum[:-1]] + [n]
def run(weights, with_d2h, two_streams=False, via_lib=False, d2h_slabs=1):
    b = bounds(weights); ts = []
    for _ in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for k in range(len(weights)):
            st = (sA, sB)[k % 2] if two_streams else sA
            if via_lib:
                lib.scrib200_h2d(ctypes.c_void_p(d_in.data_ptr() + b[k]), ctypes.c_void_p(pinned.data_ptr() + b[k]), b[k + 1] - b[k], ctypes.c_void_p(st.cuda_stream))
            else:
                with torch.cuda.stream(st):
                    d_in[b[k]:b[k + 1]].copy_(pinned[b[k]:b[k + 1]], non_blocking=True)
        if with_d2h:
            step = n // d2h_slabs
            with torch.cuda.stream(s2):
                for k in range(d2h_slabs):
                    dst[k * step:(k + 1) * step].copy_(d_out[k * step:(k + 1) * step], non_blocking=True)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts[1:])
W = {"1": (1,), "3": (1, 2, 1), "5": (1, 3, 4, 3, 1), "9": (1, 2, 3, 4, 4, 4, 3, 2, 1), "9 equal": (1,) * 9, "24 equal": (1,) * 24}
print("slabs | alone | with d2h | with d2h in 8 | two streams, with d2h | via scrib200_h2d, with d2h")
for name, w in W.items():
    print(f"{name:9s} {run(w, False):.2f}  {run(w, True):.2f}  {run(w, True, d2h_slabs=8):.2f}  {run(w, True, two_streams=True):.2f}  {run(w, True, via_lib=True):.2f}")
