"""Developer script: small calls of the round-2 kernels (DMMA rotation, time-lane expectation, warp multishuffle, tile scan,
captured transform) for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_b200 import ops, flux, utilities as ut
from scri_inputs import smooth_modes
rng = np.random.default_rng(0)
for lmin, lmax, n in ((2, 8, 77), (2, 16, 45), (0, 16, 33), (3, 11, 50), (13, 16, 40)):
    t, data = smooth_modes(n_times=n, ell_min=lmin, ell_max=lmax, seed=lmax)
    Rs = rng.normal(size=(n, 4)); Rs /= np.linalg.norm(Rs, axis=1)[:, None]
    Rs[3] = [1, 0, 0, 0]
    print("rotate", lmin, lmax, flush=True)
    ops.rotate_modes(data.copy(), Rs, lmin, lmax)
    torch.cuda.synchronize()
for lmax in (8, 16):
    n = lmax * (lmax + 2) - 3
    a = rng.normal(size=(70, n)) + 1j * rng.normal(size=(70, n))
    b = rng.normal(size=(70, n)) + 1j * rng.normal(size=(70, n))
    mats = [flux.p_plus(2, lmax, s=-2), flux.p_minus(2, lmax, s=-2), flux.p_z(2, lmax, s=-2)]
    print("expectation", lmax, flush=True)
    ops.sparse_expectation(a, a, mats); ops.sparse_expectation(a, b, mats)
    torch.cuda.synchronize()
for bw in (32, 64):
    for n in (1, 37, 300, 5000):
        x = rng.integers(0, 2**bw, size=n, dtype=np.dtype(f"u{bw // 8}"), endpoint=False)
        for widths in ((1,) * bw, (8,) * (bw // 8), (bw,), tuple([3, 5] + [1] * (bw - 8))):
            ut.multishuffle(widths)(x)
    print("multishuffle", bw, flush=True)
    torch.cuda.synchronize()
t, data = smooth_modes(n_times=2000, t0=0.0, t1=200.0, seed=3)
for order in (1, 2):
    print("antiderivative", order, flush=True)
    ops.spline_calculus(t, data, "antiderivative", order)
    torch.cuda.synchronize()
print("done")
