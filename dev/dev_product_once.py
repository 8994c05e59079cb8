"""One launch of the fused product kernel at config-4 size (for ncu).  Dev tool (GPU)."""
import sys

import torch

sys.path.insert(0, ".")
from scri_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
L = 32
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.view_as_complex(torch.randn((N, (L + 1) ** 2, 2), dtype=torch.float64, device="cuda", generator=g))
b = torch.view_as_complex(torch.randn((N, (L + 1) ** 2, 2), dtype=torch.float64, device="cuda", generator=g))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
shape = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ops._product_device_tables((2, 0, L, -2, 0, L, 129, 129, 32, shape))[0].cfg[14] = skip
for _ in range(2):
    out = ops.modes_product(a, 2, 0, L, b, -2, 0, L, 129, 129, 32, shape=shape)
torch.cuda.synchronize()
print(out.abs().max().item())
