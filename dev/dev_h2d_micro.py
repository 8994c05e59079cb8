import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
from scri_b200 import ops
w = bench.make_workload(100_000)
a = w.data
print("base type", type(a.base), "owndata", a.flags["OWNDATA"], "aligned", a.flags["ALIGNED"], "addr%64", a.ctypes.data % 64, "pinned?", torch.from_numpy(a).is_pinned(), "threads", torch.get_num_threads())
b = np.array(a)  # fresh copy in malloc'd memory
c = np.random.default_rng(0).standard_normal((100000, 154)).view(np.complex128)
def timeit(f, n=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for name, x in (("w.data", a), ("np.array(w.data)", b), ("fresh random", c)):
    print(f"{name:18s} to_device {timeit(lambda: ops.to_device(x)):.2f} ms   pinned? {torch.from_numpy(x).is_pinned()}")
stage = torch.empty(16 << 20, dtype=torch.uint8, pin_memory=True)
for name, x in (("w.data", a), ("np.array(w.data)", b)):
    src = torch.from_numpy(x).reshape(-1).view(torch.uint8)
    print(f"{name:18s} cpu copy of 16 MiB into pinned: {timeit(lambda: stage.copy_(src[:16 << 20])):.2f} ms; into pageable: {timeit(lambda: torch.empty(16 << 20, dtype=torch.uint8).copy_(src[:16 << 20])):.2f} ms")
# direct non_blocking copy from a pinned source
src = torch.from_numpy(a)
if src.is_pinned():
    print("direct H2D from pinned source", timeit(lambda: src.to("cuda", non_blocking=True)), "ms")
