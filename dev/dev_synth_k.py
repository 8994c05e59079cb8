"""Developer script: synthesis TFLOP/s against the contraction length (ell_max): separates per-tile overhead from the main loop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_b200 import ops, plan as P
from scri_inputs import real_supertranslation, smooth_modes
kw = dict(supertranslation=real_supertranslation(4), frame_rotation=[1.0, 2.0, 3.0, 4.0], boost_velocity=[0.01, 0.02, 0.03])
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for L, N in ((4, 200000), (8, 100000), (12, 60000), (16, 40000), (24, 20000)):
    n = L * (L + 2) - 3
    _, data = smooth_modes(n_times=N, ell_max=L, t0=0.0, t1=1e3)
    pl = P.TransformPlan(2, L, sb.h, **kw)
    ad = ops.to_device(data)
    ts = []
    for it in range(4):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); F = pl.synthesize(ad); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    print(f"ell_max {L:2d}: K = {2*n:4d}, G = {pl.G:5d}, N = {N}: {ms:.3f} ms, {8.0*n*pl.G*N/ms/1e9:.1f} TFLOP/s, output {16.0*pl.G*N/1e9:.2f} GB ({16.0*pl.G*N/ms/1e6:.0f} GB/s)")
    del F, ad, pl
