import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
if len(sys.argv) > 1:
    from scri_b200 import ops
    a = np.random.default_rng(0).standard_normal((100000, 154)).view(np.complex128)
    ops.to_device(a); torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter(); ops.to_device(a); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print("threads", sys.argv[1], "to_device %.2f ms" % (1e3 * min(ts)))
    if sys.argv[1] == "16":
        rt = torch.cuda.cudart()
        t0 = time.perf_counter(); r = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0); t1 = time.perf_counter()
        src = torch.from_numpy(a); dst = torch.empty_like(src, device="cuda")
        torch.cuda.synchronize(); t2 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize(); t3 = time.perf_counter()
        rt.cudaHostUnregister(a.ctypes.data); t4 = time.perf_counter()
        print("cudaHostRegister %s: %.2f ms, copy %.2f ms, unregister %.2f ms" % (r, 1e3 * (t1 - t0), 1e3 * (t3 - t2), 1e3 * (t4 - t3)))
else:
    print("cpu count", os.cpu_count())
    for n in (2, 4, 8, 12, 16):
        env = dict(os.environ, SCRIB200_COPY_THREADS=str(n))
        print(subprocess.run([sys.executable, __file__, str(n)], env=env, capture_output=True, text=True).stdout.strip())
