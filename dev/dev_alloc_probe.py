import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
kw = bench.transformation_kwargs()
w = bench.make_workload(100_000)
def stats():
    s = torch.cuda.memory_stats()
    return s.get("num_device_alloc", 0), s.get("num_device_free", 0), s["reserved_bytes.all.current"] >> 20
for it in range(8):
    s0 = stats(); torch.cuda.synchronize(); t0 = time.perf_counter()
    out = w.transform(**kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0; s1 = stats()
    print(f"call {it}: {1e3*dt:.2f} ms  cudaMalloc +{s1[0]-s0[0]} cudaFree +{s1[1]-s0[1]} reserved {s1[2]} MiB")
hs = torch._C._host_emptyCache if hasattr(torch._C, "_host_emptyCache") else None
t0 = time.perf_counter(); x = torch.empty(120_000_000, dtype=torch.uint8, pin_memory=True); print("fresh pinned alloc 120MB: %.2f ms" % (1e3*(time.perf_counter()-t0)))
del x
t0 = time.perf_counter(); x = torch.empty(120_000_000, dtype=torch.uint8, pin_memory=True); print("second pinned alloc 120MB: %.2f ms" % (1e3*(time.perf_counter()-t0)))
t0 = time.perf_counter(); y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); torch.cuda.synchronize(); print("cudaMalloc-ish 1GiB: %.2f ms" % (1e3*(time.perf_counter()-t0)))
