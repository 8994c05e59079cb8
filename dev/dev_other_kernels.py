"""Developer script: device-resident timings of the HBM-bound kernels at config-5 size (l<=16, 1e6 steps: rotation,
mode calculations, fluxes, spline calculus) against the measured HBM peak.  Prints a markdown table."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_b200 import ops, flux, _lib
N, LMIN, LMAX = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 2, int(sys.argv[2]) if len(sys.argv) > 2 else 16
n = LMAX * (LMAX + 2) - LMIN**2 + 1
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
hbm = peaks["hbm_gbs"]
g = torch.Generator(device="cuda").manual_seed(0)
t = torch.linspace(0.0, 1e5, N, dtype=torch.float64, device="cuda")
w = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 0.45 + 0.05
c = torch.randn(n, dtype=torch.complex128, device="cuda", generator=g)
data = c[None, :] * torch.exp(1j * w[None, :] * t[:, None])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(f, reps=4):
    f(); ts = []
    for i in range(reps):
        flush.fill_(i); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), r
rows = []
def report(name, ms, bytes_, note=""):
    gbs = bytes_ / (ms * 1e-3) / 1e9
    rows.append(f"| {name} | {ms:.3f} | {bytes_/1e9:.2f} | {gbs:.0f} | {100*gbs/hbm:.0f} % | {note} |")
B = 16.0 * n * N
ms, ddot = timeit(lambda: ops.spline_calculus(t, data, "derivative", 1)); report("K8 data_dot (spline_prepare + spline_tile<1>)", ms, 2 * B, "read a, write a-dot")
ms, _ = timeit(lambda: ops.spline_calculus(t, data, "antiderivative", 1)); report("K8 data_int (spline_tile<3> with in-tile sums + tile totals + add pass)", ms, 2 * B, "one extra pass over the output")
ms, _ = timeit(lambda: ops.norm(data)); report("norm", ms, B + 8 * N)
ms, (LL, Ldt) = timeit(lambda: ops.ll_ldt(data, ddot, LMIN, LMAX)); report("K5 <LL> + <L d/dt>", ms, 2 * B + 96 * N, "reads a and a-dot")
ms, _ = timeit(lambda: ops.ll_ldt(data, None, LMIN, LMAX)); report("K5 <LL> only", ms, B + 72 * N)
rd = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64, device="cuda")
ms, dpa = timeit(lambda: ops.dominant_eigenvector(LL, np.array([0.0, 0.0, 1.0]), 0)); report("K6 dominant eigenvector (Jacobi + sign scan)", ms, (72 + 24) * N)
ms, om = timeit(lambda: ops.solve3(LL, Ldt, -1.0)); report("solve3 (angular velocity)", ms, (72 + 24 + 24) * N)
mats = [flux.p_plus(LMIN, LMAX, s=-2), flux.p_minus(LMIN, LMAX, s=-2), flux.p_z(LMIN, LMAX, s=-2)]
ms, _ = timeit(lambda: ops.sparse_expectation(ddot, ddot, mats)); report("K7 momentum flux (3 matrices, one pass, time-lane kernel)", ms, B + 48 * N, "a = b = a-dot")
jm = [flux.j_plus(LMIN, LMAX), flux.j_minus(LMIN, LMAX), flux.j_z(LMIN, LMAX)]
ms, _ = timeit(lambda: ops.sparse_expectation(ddot, data, jm)); report("K7 angular-momentum flux (3 matrices)", ms, 2 * B + 48 * N)
q = torch.randn(N, 4, dtype=torch.float64, device="cuda", generator=g); q = q / q.norm(dim=1, keepdim=True)
sp = torch.complex(q[:, 0], q[:, 3]), torch.complex(q[:, 2], q[:, 1])
spin = torch.stack(sp, dim=1).contiguous()
d2 = data.clone()
ms, _ = timeit(lambda: ops.rotate_modes(d2, spin, LMIN, LMAX)); report("K4 rotation by a rotor series (in place, DMMA)", ms, 2 * B + 32 * N)
print(f"\ninputs: N = {N}, ell = {LMIN}..{LMAX} (n = {n}), modes {B/1e9:.2f} GB; HBM peak (MEASURED_PEAKS.json) {hbm:.0f} GB/s; kernels launched {_lib.launch_count()}\n")
print("| kernel | ms | algorithmic GB | GB/s | of HBM peak | note |\n|---|---|---|---|---|---|")
print("\n".join(rows))
