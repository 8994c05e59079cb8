"""Timing of the fused separable product (K9) at BASELINE config 4 size, against the dense chain.  Dev tool (GPU)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from scri_b200 import _product, ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
L = 32
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn((N, (L + 1) ** 2, 2), dtype=torch.float64, device="cuda", generator=g)
b = torch.randn((N, (L + 1) ** 2, 2), dtype=torch.float64, device="cuda", generator=g)
a[:, :4] = 0
b[:, :4] = 0
a = torch.view_as_complex(a)
b = torch.view_as_complex(b)
tb = _product.product_tables(2, 0, L, -2, 0, L, 129, 129, 32)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for shape in (0, 1, 2):
    ms = timed(lambda: ops.modes_product(a, 2, 0, L, b, -2, 0, L, 129, 129, 32, shape=shape))
    print(f"fused N={N} shape={shape}: {ms:.2f} ms  {tb.flops_per_step * N / ms / 1e9:.2f} TFLOP/s algorithmic  "
          f"{16 * 3 * 1089 * N / ms / 1e6:.0f} GB/s algorithmic", flush=True)
for shape in (0, 2):
    tbs = ops._product_device_tables((2, 0, L, -2, 0, L, 129, 129, 32, shape))[0]
    for skip in (1, 2, 4, 3, 5, 6, 7):
        tbs.cfg[14] = skip
        ms = timed(lambda: ops.modes_product(a, 2, 0, L, b, -2, 0, L, 129, 129, 32, shape=shape))
        print(f"  shape {shape} skipping stages {[x for i, x in enumerate('ABC') if skip >> i & 1]}: {ms:.2f} ms")
    tbs.cfg[14] = 0
Nd = min(N, 2000)
ms = timed(lambda: ops.grid_multiply(a[:Nd], 2, 0, L, b[:Nd], -2, 0, L, 129, 129, 64, output_ell_max=32, fused=False), reps=1)
print(f"dense chain N={Nd}: {ms:.2f} ms -> {ms * N / Nd:.1f} ms scaled to N={N}")
