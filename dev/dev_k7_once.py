"""Dev tool (GPU): the momentum-flux expectation kernel alone (for ncu): python dev/dev_k7_once.py [N] [ell_max]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scri_b200 import ops, flux
N, LMIN, LMAX = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000, 2, int(sys.argv[2]) if len(sys.argv) > 2 else 16
n = LMAX * (LMAX + 2) - LMIN**2 + 1
data = torch.randn(N, n, dtype=torch.complex128, device="cuda")
mats = [flux.p_plus(LMIN, LMAX, s=-2), flux.p_minus(LMIN, LMAX, s=-2), flux.p_z(LMIN, LMAX, s=-2)]
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.sparse_expectation(data, data, mats); e1.record(); torch.cuda.synchronize()
    print(f"momentum flux N={N} ell<={LMAX}: {e0.elapsed_time(e1):.3f} ms")
