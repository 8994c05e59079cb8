"""Developer script: run single stages of the transform at config-2 size (for ncu / quick timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scri_b200 as sb
from scri_b200 import ops, plan as P
from scri_inputs import real_supertranslation, smooth_modes
stage = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = 100000
t = np.linspace(0, 1e4, N)
_, data = smooth_modes(n_times=N, t0=0.0, t1=1e4)
kw = dict(supertranslation=real_supertranslation(4), frame_rotation=[1.0, 2.0, 3.0, 4.0], boost_velocity=[0.01, 0.02, 0.03])
pl = P.TransformPlan(2, 8, sb.h, **kw)
td = ops.to_device(t); ad = ops.to_device(data)
F = pl.synthesize(ad); up = pl.output_times(td); grid = pl.remap(td, F, up); m = pl.analyze(grid)
gridT = pl.remap_tiled(td, F, up); m2 = pl.analyze_tiled(gridT, up.shape[0])
torch.cuda.synchronize()
print("tiled vs time-major: grid", float((gridT.permute(0, 2, 1).reshape(-1, pl.G)[:up.shape[0]] - grid).abs().max()), "modes", float((m2 - m).abs().max()), "max|m|", float(m.abs().max()))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for it in range(reps):
    flush.fill_(it)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    e[0].record()
    if stage in ("all", "synth"): F = pl.synthesize(ad)
    e[1].record()
    if stage in ("all", "remap"): grid = pl.remap(td, F, up)
    e[2].record()
    if stage in ("all", "analysis"): m = pl.analyze(grid)
    e[3].record()
    if stage in ("all", "remap", "gmajor"): gridT = pl.remap_tiled(td, F, up)
    e[4].record()
    if stage in ("all", "analysis", "gmajor"): m2 = pl.analyze_tiled(gridT, up.shape[0])
    e[5].record()
    torch.cuda.synchronize()
    print("synth %.3f | remap %.3f analysis %.3f | remap_g %.3f analysis_g %.3f ms" % tuple(e[i].elapsed_time(e[i + 1]) for i in range(5)))
