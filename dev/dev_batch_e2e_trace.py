"""Developer script: where the time of parallel.transform_batch_host goes (configs[2] shard of 4096 waveforms through one
pinned block of 512): per sub-batch H2D / kernels / D2H durations and the start of each, from CUDA events."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from scri_b200 import _lib, ops, plan as P, parallel
import scri_b200 as sb
kw = bench.transformation_kwargs()
B, N, n, sub = 4096, 2048, 77, int(sys.argv[1]) if len(sys.argv) > 1 else 128
plan = P.TransformPlan(2, 8, sb.h, r_is_scaled_out=True, **kw)
t = np.linspace(0.0, 204.7, N)
host_in = torch.empty((512, N, n), dtype=torch.complex128, pin_memory=True)
host_in.numpy()[...] = (np.random.default_rng(0).normal(size=(1, N, n)) + 0j)
u, m = plan.run_batch(ops.to_device(t), ops.to_device(host_in.numpy()[:8]))
n_out = u.shape[0]
host_out = torch.empty((512, n_out, n), dtype=torch.complex128, pin_memory=True)
blocks = [host_in.numpy()] * 8
for _ in range(2):
    parallel.transform_batch_host(plan, t, blocks, sub_batch=sub, out=[host_out] * 8)
torch.cuda.synchronize(); t0 = time.perf_counter()
parallel.transform_batch_host(plan, t, blocks, sub_batch=sub, out=[host_out] * 8)
torch.cuda.synchronize(); print(f"transform_batch_host sub_batch={sub}: {(time.perf_counter() - t0) * 1e3:.1f} ms")
# the three legs alone
t_d = ops.to_device(t)
buf = torch.empty((sub, N, n), dtype=torch.complex128, device="cuda")
def timed(f, reps=8):
    f(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
h = timed(lambda: buf.copy_(host_in[:sub], non_blocking=True))
c = timed(lambda: plan.run_batch(t_d, buf))
u, m = plan.run_batch(t_d, buf)
d = timed(lambda: host_out[:sub].copy_(m, non_blocking=True))
print(f"per sub-batch alone: h2d {h:.2f} ms, kernels {c:.2f} ms, d2h {d:.2f} ms; x {B // sub} sub-batches: {B // sub * h:.0f} / {B // sub * c:.0f} / {B // sub * d:.0f} ms")
# host time of one run_batch call (launch overhead)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(8): plan.run_batch(t_d, buf)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"host time per run_batch call: {(t1 - t0) / 8 * 1e3:.2f} ms")
