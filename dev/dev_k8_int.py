"""Developer script: one antiderivative (order -1, then -2) of a 1e6 x 285 series, for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scri_b200 import ops
N, n = 1_000_000, 285
t = torch.linspace(0.0, 1e5, N, dtype=torch.float64, device="cuda")
data = torch.randn(N, n, dtype=torch.complex128, device="cuda")
for order in (1, 2):
    ops.spline_calculus(t, data, "antiderivative", order)
    torch.cuda.synchronize()
