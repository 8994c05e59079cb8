"""Developer script (not a test): print GPU-vs-oracle errors stage by stage.  Run on the GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import scri_b200 as sb
from scri_b200 import ops, plan as P, _sf
from oracle import scri_ref as R, quat, sf as osf, spinsfast as ospf
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from scri_inputs import real_supertranslation, smooth_modes, rotor_set

def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

print("device", torch.cuda.get_device_name(0))
# ---- rotation
t, data = smooth_modes(n_times=300)
rng = np.random.default_rng(1)
Rs = quat.normalized(rng.normal(size=(300, 4)))
d1 = data.copy(); ops.rotate_modes(d1, Rs, 2, 8)
Wo = R.Modes(t=t, data=data.copy()); R.rotate_decomposition_basis(Wo, Rs)
print("rotate series rel err", rel(d1, Wo.data))
d1 = data.copy(); ops.rotate_modes(d1, Rs[0], 2, 8)
Wo = R.Modes(t=t, data=data.copy()); R.rotate_decomposition_basis(Wo, Rs[0])
print("rotate const  rel err", rel(d1, Wo.data))
d1 = data.copy(); ops.rotate_modes(d1, np.array([1.0, 0, 0, 0]), 2, 8)
print("identity bit-exact", np.array_equal(d1, data))

# ---- transform stage by stage
st = real_supertranslation(4)
kw = dict(supertranslation=st, frame_rotation=[1.0, 2.0, 3.0, 4.0], boost_velocity=[0.01, 0.02, 0.03])
t, data = smooth_modes(n_times=801, t0=0.0, t1=80.0)
Wo = R.Modes(t=t, data=data.copy())
g_o, inter = R.from_modes(Wo, return_intermediates=True, **kw)
pl = P.TransformPlan(2, 8, sb.h, **kw)
print("k err", rel(pl.kconformal, inter["kconformal"].ravel()), "alpha err", rel(pl.alpha, inter["alpha"].ravel()))
td = ops.to_device(t); ad = ops.to_device(data)
F = pl.synthesize(ad)
print("synth rel err", rel(F.cpu().numpy(), inter["synthesized"].reshape(len(t), -1)))
up = pl.output_times(td)
print("n_out", up.shape[0], g_o.t.shape[0], "uprm err", abs(up.cpu().numpy() - g_o.t).max() if up.shape[0] == g_o.t.shape[0] else None)
grid = pl.remap(td, F, up)
print("remap rel err", rel(grid.cpu().numpy(), g_o.data))
m_o = R.to_modes(g_o, 8)
m_g = pl.analyze(grid)
print("analysis rel err", rel(m_g.cpu().numpy(), m_o.data))
# analysis alone on oracle grid
m_g2 = ops.map2salm(g_o.data.reshape(-1, g_o.n_theta, g_o.n_phi), -2, 8)[:, 4:]
print("analysis(alone) rel err", rel(m_g2, m_o.data))

# small chunk to exercise halos
pl.spline_body = 100
grid2 = pl.remap(td, F, up)
print("remap chunk=100 rel err", rel(grid2.cpu().numpy(), g_o.data), "vs chunk default", rel(grid2.cpu().numpy(), grid.cpu().numpy()))

# nonuniform times
t, data = smooth_modes(n_times=500, uniform=False)
Wo = R.Modes(t=t, data=data.copy())
m_o = R.transform(Wo, **kw)
w = sb.WaveformModes(t=t, data=data.copy(), ell_min=2, ell_max=8, frameType=sb.Inertial, dataType=sb.h, r_is_scaled_out=True, m_is_scaled_out=True)
m_g = w.transform(**kw)
print("full transform nonuniform: n", m_g.n_times, m_o.t.shape[0], "rel err", rel(m_g.data, m_o.data) if m_g.n_times == m_o.t.shape[0] else None)

# ---- spline derivative
from scipy.interpolate import CubicSpline
dd = ops.spline_calculus(t, data, "derivative", 1)
print("data_dot rel err", rel(dd, CubicSpline(t, data).derivative()(t)))
dd2 = ops.spline_calculus(t, data, "derivative", 2)
print("data_ddot rel err", rel(dd2, CubicSpline(t, data).derivative(2)(t)))
tp = np.linspace(t[0], t[-1], 777)
print("interp rel err", rel(ops.spline_calculus(t, data, "evaluate", tprime=tp), CubicSpline(t, data)(tp)))

# ---- mode calculations
Wo = R.Modes(t=t, data=data.copy())
LL, Ldt = ops.ll_ldt(data, dd, 2, 8)
print("LL rel err", rel(LL, R.LLMatrix(Wo)), "Ldt rel err", rel(Ldt, R.LdtVector(Wo)))
print("LVector rel err", rel(ops.l_vector(data, data, 2, 8), R.LVector(Wo)))
print("norm rel err", rel(ops.norm(data), R.norm(Wo)))
w = sb.WaveformModes(t=t, data=data.copy(), ell_min=2, ell_max=8, frameType=sb.Inertial, dataType=sb.h, r_is_scaled_out=True, m_is_scaled_out=True)
print("dpa err", abs(w.LLDominantEigenvector() - R.LLDominantEigenvector(Wo)).max())
print("omega rel err", rel(w.angular_velocity(), R.angular_velocity(Wo)))
print("Edot rel err", rel(w.energy_flux(), R.energy_flux(Wo)))
print("pdot rel err", rel(w.momentum_flux(), R.momentum_flux(Wo)))
print("jdot rel err", rel(w.angular_momentum_flux(), R.angular_momentum_flux(Wo)))

# ---- timing at config 2 size
N = 100000
t = np.linspace(0, 1e4, N)
_, data = smooth_modes(n_times=N, t0=0.0, t1=1e4)
td = ops.to_device(t); ad = ops.to_device(data)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record(); F = pl.synthesize(ad); ev[1].record(); up = pl.output_times(td); ev[2].record()
    pl.spline_body = 0
    grid = pl.remap(td, F, up); ev[3].record(); m = pl.analyze(grid); ev[4].record()
    torch.cuda.synchronize()
    print("N=1e5: synth %.3f ms, times %.3f ms, remap %.3f ms, analysis %.3f ms, wall %.3f ms" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[3].elapsed_time(ev[4]), (time.time() - t0) * 1e3))
