"""Restatement of spinsfast.salm2map / map2salm (Huffenberger & Wandelt 2010, ApJS 189:255).

TEST INFRASTRUCTURE (see oracle/__init__.py).  `spinsfast>=2022.4` (C + FFTW) is not vendored
under /root/reference.  Call sites followed: scri/waveform_grid.py:303-307 (map2salm per time
step), scri/modes_time_series.py:177-188, scri/asymptotic_bondi_data/transformations.py:419-429,
tests/test_flux.py:38-113.

Grid: theta_j = pi j/(Ntheta-1) (both poles), phi_k = 2 pi k/Nphi;
salm index l(l+1)+m from l=0 (entries with l<|s| are zero).

map2salm follows the published algorithm literally, with FFTs:
  1. extend the map to the 2-torus (2Ntheta-2 rings), F(2pi-theta, phi+pi) = (-1)^s F(theta,phi);
  2. multiply by the real-space quadrature weights W(theta) = IFFT of w(p) = int_0^pi sin(t) e^{ipt} dt
     (p taken over the FFT frequencies of the extended ring; the Nyquist frequency appears once);
  3. 2-D FFT -> I_{m'm};
  4. a_lm = (-1)^s sqrt((2l+1)/4pi) * sum_{m'} c^{l}_{m'}(m,-s) I_{m' m}, where
     d^l_{m,-s}(theta) = sum_{m'} c_{m'} e^{i m' theta} is the Fourier series of Wigner's d
     (H&W write c through Delta^l = d^l(pi/2)).
PARITY UNPINNED for maps with power above l = Ntheta-2 (aliasing / Nyquist handling cannot be
compared with the C library here); exact, and verified by round trip, below that.
"""
import math
from functools import lru_cache

import numpy as np

from . import sf


def salm2map(salm, s, lmax, Ntheta, Nphi):
    """f[..., j, k] = sum_lm salm[..., lm] sYlm(theta_j, phi_k)"""
    salm = np.asarray(salm, dtype=complex)
    theta = np.linspace(0.0, np.pi, Ntheta)
    phi = np.linspace(0.0, 2 * np.pi, Nphi, endpoint=False)
    d = _d_table(s, lmax, Ntheta)  # [lm, j]
    ms = np.array([m for ell in range(lmax + 1) for m in range(-ell, ell + 1)])
    E = np.exp(1j * ms[:, None] * phi[None, :])  # [lm, k]
    Y = d[:, :, None] * E[:, None, :]  # [lm, j, k]
    return np.tensordot(salm, Y, axes=([-1], [0]))


@lru_cache(maxsize=None)
def _d_table(s, lmax, Ntheta, extended=False):
    """(-1)^s sqrt((2l+1)/4pi) d^l_{m,-s}(theta_j) for all lm (zeros for l<|s|).  [lm, j]"""
    n = 2 * Ntheta - 2 if extended else Ntheta
    theta = np.pi * np.arange(n) / (Ntheta - 1)
    out = np.zeros(((lmax + 1) ** 2, n))
    for ell in range(abs(s), lmax + 1):
        f = (-1) ** s * math.sqrt((2 * ell + 1) / (4 * math.pi))
        for m in range(-ell, ell + 1):
            out[sf.LM_index(ell, m, 0)] = f * np.asarray(sf.wigner_d_small(theta, ell, m, -s), dtype=float)
    return out


def _wquad(p):
    """int_0^pi sin(theta) e^{i p theta} dtheta"""
    if p == 1:
        return 1j * math.pi / 2
    if p == -1:
        return -1j * math.pi / 2
    if p % 2 == 0:
        return 2.0 / (1.0 - p * p)
    return 0.0


@lru_cache(maxsize=None)
def _d_fourier(s, lmax, Ntheta):
    """c[lm, m'] Fourier coefficients of (-1)^s sqrt((2l+1)/4pi) d^l_{m,-s}(theta), m' in [-lmax, lmax].

    Computed by an FFT of d sampled on a ring fine enough to hold e^{i m' theta}, |m'|<=lmax exactly.
    """
    nfine = 2 * (2 * lmax + 2)
    theta = 2 * np.pi * np.arange(nfine) / nfine
    c = np.zeros(((lmax + 1) ** 2, 2 * lmax + 1), dtype=complex)
    for ell in range(abs(s), lmax + 1):
        f = (-1) ** s * math.sqrt((2 * ell + 1) / (4 * math.pi))
        for m in range(-ell, ell + 1):
            d = f * np.asarray(sf.wigner_d_small(theta, ell, m, -s), dtype=float)
            spec = np.fft.fft(d) / nfine
            for mp in range(-ell, ell + 1):
                c[sf.LM_index(ell, m, 0), mp + lmax] = spec[mp % nfine]
    return c


def map2salm(fmap, s, lmax):
    """Huffenberger-Wandelt analysis.  fmap[..., Ntheta, Nphi] complex -> [..., (lmax+1)^2]."""
    fmap = np.asarray(fmap, dtype=complex)
    Ntheta, Nphi = fmap.shape[-2:]
    lead = fmap.shape[:-2]
    f = fmap.reshape((-1, Ntheta, Nphi))
    NG = 2 * Ntheta - 2
    # step 1: extension to the torus: F(theta_j, phi_k), j = Ntheta..NG-1 <- (-1)^s f(theta_{NG-j}, phi_k + pi)
    # done in phi-Fourier space so that odd Nphi needs no phi interpolation:
    fm = np.fft.fft(f, axis=-1) / Nphi  # [b, j, mfreq]
    mfreq = np.fft.fftfreq(Nphi, 1.0 / Nphi).astype(int)
    ext = np.empty((f.shape[0], NG, Nphi), dtype=complex)
    ext[:, :Ntheta] = fm
    sign = (-1.0) ** (np.abs(mfreq + s) % 2)
    ext[:, Ntheta:] = fm[:, Ntheta - 2 : 0 : -1] * sign[None, None, :]
    # step 2: real-space weights from the inverse FFT of w(p) over the ring's FFT frequencies
    w = np.zeros(NG, dtype=complex)
    for ip in range(NG):
        p = ip if ip <= Ntheta - 1 else ip - NG
        w[ip] = _wquad(p)
    # W(theta_j) = sum_p w(p) e^{-i p theta_j}
    Wr = (np.fft.fft(w)).real
    # step 3: theta FFT
    I = np.fft.fft(ext * Wr[None, :, None], axis=1) / NG  # I[b, m'freq, mfreq] = (1/NG) sum_j W F e^{-i m' theta_j}
    # step 4: a_lm = 2 pi sum_{m'} c_{m'} I_{-m', m}   (int sin(t) d(t) F_m(t) dt with d = sum c e^{i m' t})
    c = _d_fourier(s, lmax, Ntheta)
    out = np.zeros((f.shape[0], (lmax + 1) ** 2), dtype=complex)
    for ell in range(abs(s), lmax + 1):
        for m in range(-ell, ell + 1):
            lm = sf.LM_index(ell, m, 0)
            acc = 0.0
            for mp in range(-ell, ell + 1):
                acc = acc + c[lm, mp + lmax] * I[:, (-mp) % NG, m % Nphi]
            out[:, lm] = 2 * math.pi * acc
    return out.reshape(lead + ((lmax + 1) ** 2,))
