"""Dev tool (GPU): where do the sporadic slow steps come from?  (1) per-step device times of the bench loop with and
without the nvidia-smi clock sampler; (2) per-call end-to-end times with the pinned-host allocator's counters."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
import scri_b200 as sb
from scri_b200 import ops
from scri_b200.plan import TransformPlan

kw = bench.transformation_kwargs()
w = bench.make_workload(100_000)
plan = TransformPlan(w.ell_min, w.ell_max, w.dataType, r_is_scaled_out=True, **kw)
t_d, a_d = ops.to_device(w.t), ops.to_device(w.data)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def loop(n):
    out = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        plan.run(t_d, a_d)
        e1.record()
        torch.cuda.synchronize()
        out.append((e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)))
    return out


loop(5)
for label, sampler in (("no sampler", False), ("nvidia-smi -lms 100", True), ("no sampler", False), ("nvidia-smi -lms 100", True)):
    s = bench.ClockSampler(0)
    if sampler:
        s.start()
        time.sleep(0.3)
    r = loop(100)
    if sampler:
        s.stop()
    dev = np.array([x[0] for x in r]); wall = np.array([x[1] for x in r])
    print(f"{label:22s} device ms: median {np.median(dev):.3f} mean {dev.mean():.3f} max {dev.max():.3f} | wall ms: median {np.median(wall):.3f} max {wall.max():.3f} | steps > 1.2 x median: {(dev > 1.2 * np.median(dev)).sum()}")


def host_stats():
    try:
        st = torch.cuda.host_memory_stats()
        return {k: st[k] for k in ("num_host_alloc", "num_host_free", "allocated_bytes.current", "segment.current") if k in st} or dict(list(st.items())[:6])
    except Exception as e:
        return {"unavailable": str(e)}


print("host allocator before:", host_stats())
out = None
times = []
for i in range(40):
    t0 = time.perf_counter()
    out = w.transform(**kw)
    torch.cuda.synchronize()
    times.append(1e3 * (time.perf_counter() - t0))
    if times[-1] > 8:
        print(f"   call {i}: {times[-1]:.1f} ms; host allocator: {host_stats()}")
print("e2e ms:", " ".join(f"{x:.1f}" for x in times))
print("host allocator after:", host_stats())
import gc
print("gc counts", gc.get_count(), "gc enabled", gc.isenabled())
