"""Dev tool (GPU): one launch each of the rotation (K4), momentum-flux expectation (K7) and spline calculus (K8) kernels at
N x 285 modes, for an `ncu --set full` capture: python dev/dev_ncu_other.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scri_b200 import ops, flux
N, LMIN, LMAX = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000, 2, 16
n = LMAX * (LMAX + 2) - LMIN**2 + 1
g = torch.Generator(device="cuda").manual_seed(0)
t = torch.linspace(0.0, 0.1 * N, N, dtype=torch.float64, device="cuda")
data = torch.randn(N, n, dtype=torch.complex128, device="cuda", generator=g)
q = torch.randn(N, 4, dtype=torch.float64, device="cuda", generator=g); q = q / q.norm(dim=1, keepdim=True)
spin = torch.stack((torch.complex(q[:, 0], q[:, 3]), torch.complex(q[:, 2], q[:, 1])), dim=1).contiguous()
mats = [flux.p_plus(LMIN, LMAX, s=-2), flux.p_minus(LMIN, LMAX, s=-2), flux.p_z(LMIN, LMAX, s=-2)]
for _ in range(2):
    ddot = ops.spline_calculus(t, data, "derivative", 1)
    ops.spline_calculus(t, data, "antiderivative", 1)
    ops.sparse_expectation(ddot, ddot, mats)
    ops.rotate_modes(data, spin, LMIN, LMAX)
    torch.cuda.synchronize()
