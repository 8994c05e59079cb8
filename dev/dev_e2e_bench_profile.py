"""Developer script: profile the e2e call exactly as bench.py makes it."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
kw = bench.transformation_kwargs()
w = bench.make_workload(100_000)
print("data flags", w.data.flags["C_CONTIGUOUS"], w.data.dtype, w.data.shape, "t", w.t.dtype, w.t.flags["C_CONTIGUOUS"])
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = w.transform(**kw)
    torch.cuda.synchronize(); print("  transform call %d: %.2f ms" % (it, (time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile(); pr.enable(); out = w.transform(**kw); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
