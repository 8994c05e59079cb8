"""One separable salm2map + map2salm round trip at l <= 32 on the 65 x 65 grid (for ncu).  Dev tool (GPU)."""
import sys
import torch
sys.path.insert(0, ".")
from scri_b200 import ops  # noqa: E402
L, N = 32, 20000
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.view_as_complex(torch.randn((N, (L + 1) ** 2, 2), dtype=torch.float64, device="cuda", generator=g))
for _ in range(2):
    grid = ops.salm2map(a, -2, L, 65, 65)
    back = ops.map2salm(grid, -2, L, 65, 65)
torch.cuda.synchronize()
print(float((back - a).abs().max()))
