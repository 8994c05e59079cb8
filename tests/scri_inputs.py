"""Deterministic input builders shared by the tests and dev scripts."""
import numpy as np


# ---- shared input builders (deterministic seeds; not hash("...") seeds as in the reference's conftest.py)
def real_supertranslation(ell_max, seed=123, scale=1e-2):
    """Random supertranslation modes satisfying the reality condition (reference tests/conftest.py:182-192)."""
    rng = np.random.default_rng(seed)
    n = (ell_max + 1) ** 2
    a = scale * (rng.uniform(size=n) - 0.5 + 1j * (rng.uniform(size=n) - 0.5))
    for ell in range(ell_max + 1):
        for m in range(ell + 1):
            i_pos = ell * (ell + 1) + m
            i_neg = ell * (ell + 1) - m
            if m == 0:
                a[i_pos] = a[i_pos].real
            else:
                a[i_neg] = (-1.0) ** m * np.conj(a[i_pos])
    return a


def rotor_set(n_random=6, seed=7):
    """A few lattice rotors + random ones (reference tests/conftest.py:173-179, trimmed)."""
    rng = np.random.default_rng(seed)
    ones = [0, -1.0, 1.0]
    rs = [np.array([w, x, y, z], dtype=float) for w in ones for x in ones for y in ones for z in ones][1:]
    rs = [r / np.linalg.norm(r) for r in rs][:: max(1, len(rs) // 10)]
    rs += [r / np.linalg.norm(r) for r in rng.normal(size=(n_random, 4))]
    return rs


def smooth_modes(n_times=601, ell_min=2, ell_max=8, t0=-10.0, t1=100.0, seed=0, uniform=True):
    rng = np.random.default_rng(seed)
    n = ell_max * (ell_max + 2) - ell_min**2 + 1
    if uniform:
        t = np.linspace(t0, t1, n_times)
    else:
        t = np.sort(rng.uniform(t0, t1, size=n_times))
        t[0], t[-1] = t0, t1
    c = rng.normal(size=n) + 1j * rng.normal(size=n)
    w = rng.uniform(0.05, 0.5, size=n)
    return t, c[None, :] * np.exp(1j * w[None, :] * t[:, None])
