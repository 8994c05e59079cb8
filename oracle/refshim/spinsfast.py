"""Stand-in for `spinsfast` (C + FFTW; Huffenberger & Wandelt 2010), built on oracle/spinsfast.py.  TEST INFRASTRUCTURE -
see oracle/refshim/quaternion/__init__.py.  Call sites: scri/waveform_grid.py:303-307, scri/modes_time_series.py:177-188,
scri/asymptotic_bondi_data/transformations.py:419-429."""
import numpy as np

from oracle import spinsfast as _sp


def N_lm(lmax):
    return (lmax + 1) ** 2


def lm_ind(ell, m, lmax=None):
    return ell * (ell + 1) + m


def ind_lm(i, lmax=None):
    ell = int(np.floor(np.sqrt(i)))
    return ell, i - ell * (ell + 1)


def salm2map(salm, s, lmax, Ntheta, Nphi):
    salm = np.asarray(salm, dtype=complex)
    return _sp.salm2map(salm, int(s), int(lmax), int(Ntheta), int(Nphi))


def map2salm(f, s, lmax):
    f = np.asarray(f, dtype=complex)
    return _sp.map2salm(f, int(s), int(lmax))
