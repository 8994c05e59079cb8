"""Developer script: H2D rate of one 124 MB block from (a) torch pinned memory, (b) a numpy array page-locked in place by
cudaHostRegister (what ops.to_device_slabs does with a caller's array), alone and with a D2H running the other way."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scri_b200 import ops, _lib
n = 124_000_000
pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True); pinned.fill_(1)
arr = np.ones(n, dtype=np.uint8)
arr2 = np.ones((100000, 78), dtype=np.complex128)[:, :77].copy(); arr2 = np.ones(n // 16, dtype=np.complex128)
print("registered:", ops._maybe_register(arr), ops._maybe_register(arr), ops._maybe_register(arr2), ops._maybe_register(arr2))
reg = torch.from_numpy(arr)
reg2 = torch.from_numpy(arr2.view(np.uint8).reshape(-1))
# a registered block whose pages are transparent huge pages
import mmap
mm = mmap.mmap(-1, n + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
try:
    mm.madvise(mmap.MADV_HUGEPAGE)
except Exception as e:
    print("madvise:", e)
huge = np.frombuffer(mm, dtype=np.uint8)[: n]; huge[:] = 1
print("registered huge:", ops._maybe_register(huge), ops._maybe_register(huge))
hug = torch.from_numpy(huge)
dst = torch.empty(n, dtype=torch.uint8, pin_memory=True); dst.fill_(2)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.ones(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(src, with_d2h, slabs=1):
    ts = []
    for _ in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            step = n // slabs
            for k in range(slabs):
                d_in[k * step:(k + 1) * step].copy_(src[k * step:(k + 1) * step], non_blocking=True)
        if with_d2h:
            with torch.cuda.stream(s2): dst.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts[1:])
for name, src in (("torch pinned", pinned), ("numpy registered", reg), ("numpy c128 registered", reg2), ("THP registered", hug)):
    print(f"{name:24s} h2d alone {run(src, False):.2f} ms   with d2h {run(src, True):.2f} ms   9 slabs with d2h {run(src, True, 9):.2f} ms")
