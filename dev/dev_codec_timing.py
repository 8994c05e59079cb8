"""Dev tool (GPU): device-resident throughput of the codec stages at config-2 size (1e5 x 77 complex128 = 123 MB)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, ".")
from scri_b200 import _lib
lib = _lib.load()
N, C = 100_000, 154
a = torch.randint(0, 2**62, (N, C), dtype=torch.int64, device="cuda")
b = torch.empty_like(a)
ws = torch.empty(lib.scrib200_xor_timeseries_workspace_bytes(N, C), dtype=torch.uint8, device="cuda")
acc = torch.zeros(2, dtype=torch.int64, device="cuda")
widths = (8, 8, 4, 4, 4, 4) + (2,) * 8 + (1,) * 16
w = (ctypes.c_int * len(widths))(*widths)
st = _lib.stream_ptr()
def T(name, f, bytes_moved):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.3f} ms  {bytes_moved / ms / 1e6:.0f} GB/s")
nb = a.numel() * 8
T("xor_timeseries", lambda: lib.scrib200_xor_timeseries(_lib.ptr(a), _lib.ptr(b), N, C, 0, _lib.ptr(ws), ws.numel(), st), 2 * nb)
T("xor_timeseries_reverse", lambda: lib.scrib200_xor_timeseries(_lib.ptr(a), _lib.ptr(b), N, C, 1, _lib.ptr(ws), ws.numel(), st), 3 * nb)
T("fletcher32", lambda: lib.scrib200_fletcher32(_lib.ptr(a), a.numel() * 4, _lib.ptr(acc), st), nb)
T("multishuffle", lambda: lib.scrib200_multishuffle(_lib.ptr(a), _lib.ptr(b), a.numel(), 64, w, len(widths), 1, st), 2 * nb)
T("multishuffle reverse", lambda: lib.scrib200_multishuffle(_lib.ptr(a), _lib.ptr(b), a.numel(), 64, w, len(widths), 0, st), 2 * nb)
