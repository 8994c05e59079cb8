"""Dev tool (GPU): the rotation kernel alone (for ncu): python dev/dev_rotate_once.py [N] [ell_max]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scri_b200 import ops
N, LMIN, LMAX = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000, 2, int(sys.argv[2]) if len(sys.argv) > 2 else 16
n = LMAX * (LMAX + 2) - LMIN**2 + 1
g = torch.Generator(device="cuda").manual_seed(0)
data = torch.randn(N, n, dtype=torch.complex128, device="cuda", generator=g)
q = torch.randn(N, 4, dtype=torch.float64, device="cuda", generator=g); q = q / q.norm(dim=1, keepdim=True)
spin = torch.stack((torch.complex(q[:, 0], q[:, 3]), torch.complex(q[:, 2], q[:, 1])), dim=1).contiguous()
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.rotate_modes(data, spin, LMIN, LMAX); e1.record(); torch.cuda.synchronize()
    print(f"rotate N={N} ell<={LMAX}: {e0.elapsed_time(e1):.3f} ms")
