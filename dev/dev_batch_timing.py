"""Developer script: BASELINE config 3 (batch of random-mode waveforms, 2048 steps each, l <= 8) through run_batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_b200 import ops, plan as P
from scri_inputs import real_supertranslation
B, N, n = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 2048, 77
kw = dict(supertranslation=real_supertranslation(4), frame_rotation=[1.0, 2.0, 3.0, 4.0], boost_velocity=[0.01, 0.02, 0.03])
pl = P.TransformPlan(2, 8, sb.h, **kw)
g = torch.Generator(device="cuda").manual_seed(0)
t = torch.linspace(0.0, 204.7, N, dtype=torch.float64, device="cuda")
w = torch.rand(B, 1, n, dtype=torch.float64, device="cuda", generator=g) * 0.45 + 0.05
c = torch.randn(B, 1, n, dtype=torch.complex128, device="cuda", generator=g)
data = c * torch.exp(1j * w * t[None, :, None])
for it in range(4):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); up, out = pl.run_batch(t, data); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"run_batch B={B}: {ms:.3f} ms -> {B * N * n / ms / 1e6:.2f} G mode-timesteps/s (n_out {up.shape[0]})")
torch.cuda.synchronize(); t0 = time.perf_counter()
for b in range(min(B, 64)): pl.run(t, data[b])
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"per-waveform loop: {1e3 * dt / min(B, 64):.3f} ms per waveform -> {N * n / (dt / min(B, 64)) / 1e9:.2f} G mode-timesteps/s")
