"""Timing of the separable salm2map against the dense synthesis at BASELINE config 4 sizes.  Dev tool (GPU)."""
import sys

import torch

sys.path.insert(0, ".")
from scri_b200 import ops  # noqa: E402

L = 32
for nth, N in ((65, 20000), (129, 8000)):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.view_as_complex(torch.randn((N, (L + 1) ** 2, 2), dtype=torch.float64, device="cuda", generator=g))
    for sep in (True, False):
        ops.salm2map(a, -2, L, nth, nth, separable=sep)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.salm2map(a, -2, L, nth, nth, separable=sep)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"salm2map l<=32 grid {nth}x{nth} N={N} {'separable' if sep else 'dense    '}: {ms:.2f} ms -> {ms * 1e5 / N:.1f} ms per 1e5 steps", flush=True)
    del a

for nth, N in ((65, 20000), (129, 8000)):
    g = torch.Generator(device="cuda").manual_seed(1)
    grid = torch.view_as_complex(torch.randn((N, nth, nth, 2), dtype=torch.float64, device="cuda", generator=g))
    ops.map2salm(grid, -2, L, nth, nth)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.map2salm(grid, -2, L, nth, nth)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"map2salm l<=32 grid {nth}x{nth} N={N}: {ms:.2f} ms -> {ms * 1e5 / N:.1f} ms per 1e5 steps", flush=True)
    del grid
