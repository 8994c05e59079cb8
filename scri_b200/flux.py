"""Energy / momentum / angular-momentum fluxes (mirrors scri/flux.py:11-441), expectation values on the GPU.

The sparse matrix elements (p_z, p_+-, j_z, j_+-) are tiny host tables built once per (ell_min, ell_max) and
cached, as in the reference (flux.py:11-37 lru_cache); the per-time-step sums <a|M|b>(t) - the hot loop
flux.py:40-78 - run as one fused pass over the modes for all three components.
"""
import functools
import math

import numpy as np

from . import _sf, ops
from .constants import h as htype
from .constants import hdot as hdottype


@functools.lru_cache(maxsize=None)
def _log_factorials(n):
    out = [0.0] * (n + 1)
    for i in range(2, n + 1):
        out[i] = out[i - 1] + math.log(i)
    return out


def clebsch_gordan(j1, m1, j2, m2, j3, m3):
    """<j1 m1 j2 m2|j3 m3> for integer arguments (sf.clebsch_gordan argument order; scri/flux.py:7)."""
    if m1 + m2 != m3 or abs(m1) > j1 or abs(m2) > j2 or abs(m3) > j3 or j3 < abs(j1 - j2) or j3 > j1 + j2:
        return 0.0
    f = math.factorial
    pre = (2 * j3 + 1) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) / f(j1 + j2 + j3 + 1)
    pre *= f(j3 + m3) * f(j3 - m3) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2)
    s = 0.0
    for k in range(max(0, j2 - j3 - m1, j1 - j3 + m2), min(j1 + j2 - j3, j1 - m1, j2 + m2) + 1):
        s += (-1) ** k / (f(k) * f(j1 + j2 - j3 - k) * f(j1 - m1 - k) * f(j2 + m2 - k) * f(j3 - j2 + m1 + k) * f(j3 - j1 - m2 + k))
    return math.sqrt(pre) * s


def _as_matrix(elements, ell_min):
    rows, cols, vals = zip(*((_sf.LM_index(lp, mp, ell_min), _sf.LM_index(l, m, ell_min), v) for lp, mp, l, m, v in elements))
    return np.array(rows, dtype=np.int32), np.array(cols, dtype=np.int32), np.array(vals, dtype=complex)


@functools.lru_cache(maxsize=None)
def p_z(ell_min, ell_max, s=-2):
    """<s,j,m|cos(theta)|s,l,m> (scri/flux.py:213-246)"""
    out = []
    for ell in range(ell_min, ell_max + 1):
        for ellp in range(max(ell_min, ell - 1), min(ell_max, ell + 1) + 1):
            for m in range(-ell, ell + 1):
                if abs(m) > ellp:
                    continue
                v = math.sqrt((2.0 * ell + 1.0) / (2.0 * ellp + 1.0)) * clebsch_gordan(ell, m, 1, 0, ellp, m) * clebsch_gordan(ell, -s, 1, 0, ellp, -s)
                out.append((ellp, m, ell, m, v))
    return _as_matrix(out, ell_min)


@functools.lru_cache(maxsize=None)
def p_plusminus(ell_min, ell_max, sign, s=-2):
    """p+- = sin(theta) exp(+-i phi) matrix elements (scri/flux.py:249-294)"""
    if sign not in (1, -1):
        raise ValueError("sign must be either 1 or -1 in p_plusminus")
    prefac = -1.0 * sign * math.sqrt(8.0 * math.pi / 3.0)
    out = []
    for ell in range(ell_min, ell_max + 1):
        for ellp in range(max(ell_min, ell - 1), min(ell_max, ell + 1) + 1):
            for m in range(-ell, ell + 1):
                mp = m + sign
                if abs(mp) > ellp:
                    continue
                el = math.sqrt(3.0 * (2.0 * ell + 1.0) / (4.0 * math.pi * (2.0 * ellp + 1))) * clebsch_gordan(1, sign, ell, m, ellp, mp) * clebsch_gordan(1, 0, ell, -s, ellp, -s)
                out.append((ellp, mp, ell, m, prefac * el))
    return _as_matrix(out, ell_min)


p_plus = functools.partial(p_plusminus, sign=+1)
p_minus = functools.partial(p_plusminus, sign=-1)


@functools.lru_cache(maxsize=None)
def j_z(ell_min, ell_max):
    """<j,n|j^z|l,m> = i m (scri/flux.py:345-358)"""
    return _as_matrix([(ell, m, ell, m, 1.0j * m) for ell in range(ell_min, ell_max + 1) for m in range(-ell, ell + 1)], ell_min)


@functools.lru_cache(maxsize=None)
def j_plusminus(ell_min, ell_max, sign):
    """<j,n|j^+-|l,m> = i sqrt((l-+m)(l+-m+1)) (scri/flux.py:361-385)"""
    if sign not in (1, -1):
        raise ValueError("sign must be either 1 or -1 in j_plusminus")
    out = []
    for ell in range(ell_min, ell_max + 1):
        for m in range(-ell, ell + 1):
            mp = m + sign
            if abs(mp) > ell:
                continue
            out.append((ell, mp, ell, m, 1.0j * math.sqrt((ell - m * sign) * (ell + m * sign + 1))))
    return _as_matrix(out, ell_min)


j_plus = functools.partial(j_plusminus, sign=+1)
j_minus = functools.partial(j_plusminus, sign=-1)


def sparse_expectation_value(abar, rows, columns, values, b):
    """<a|M|b> given conj(a) - same contract as scri/flux.py:40-78."""
    return ops.sparse_expectation(np.conj(abar), b, [(rows, columns, values)])[:, 0]


def intersection(t1, t2, min_step=None, min_time=None, max_time=None):
    """Common time axis of two sequences (scri/extrapolation.py:47-125): it starts at the later of the two first samples
    (or `min_time`), ends before the earlier of the two last samples (or `max_time`), and at each point steps by the smaller
    of the two local sample spacings, but at least `min_step` (default: the smallest spacing in either input)."""
    t1 = np.asarray(t1, dtype=float)
    t2 = np.asarray(t2, dtype=float)
    if t1.size == 0:
        raise ValueError("t1 is empty.  Assuming this is not desired.")
    if t2.size == 0:
        raise ValueError("t2 is empty.  Assuming this is not desired.")
    min1, min2, max1, max2 = t1[0], t2[0], t1[-1], t2[-1]
    mint = max(min1, min2) if min_time is None else max(max(min1, min2), min_time)
    maxt = min(max1, max2) if max_time is None else min(min(max1, max2), max_time)
    if mint > max1 or mint > max2:
        raise ValueError(f"Empty intersection in t1=[{min1}, ..., {max1}], t2=[{min2}, ..., {max2}] with min_time={min_time}")
    if maxt < min1 or maxt < min2:
        raise ValueError(f"Empty intersection in t1=[{min1}, ..., {max1}], t2=[{min2}, ..., {max2}] with max_time={max_time}")
    if min_step is None:
        min_step = min(np.min(np.diff(t1)), np.min(np.diff(t2)))
    t = np.empty(t1.size + t2.size)
    t[0] = mint
    i = i1 = i2 = 0
    n1, n2 = t1.size, t2.size
    while t[i] < maxt:
        # i1: t[i] lies in (t1[i1-1], t1[i1]]  (0 when t[i] is outside t1); the same for i2
        if t[i] < min1 or t[i] > max1:
            i1 = 0
        else:
            i1 = max(i1, 1)
            while t[i] > t1[i1] and i1 < n1:
                i1 += 1
        if t[i] < min2 or t[i] > max2:
            i2 = 0
        else:
            i2 = max(i2, 1)
            while t[i] > t2[i2] and i2 < n2:
                i2 += 1
        t[i + 1] = t[i] + max(min(t1[i1] - t1[i1 - 1], t2[i2] - t2[i2 - 1]), min_step)
        i += 1
        if t[i] > maxt:
            break
    return t[:i]


def matrix_expectation_value(a, M, b, allow_LM_differ=False, allow_times_differ=False):
    """(times, <a|M|b>(u)) (scri/flux.py:81-179).  `M(ell_min, ell_max)` returns (rows, columns, values).  With
    `allow_LM_differ` the product runs over the ell range the two waveforms share; with `allow_times_differ` both are
    interpolated onto the `intersection` of their time axes first (cubic splines on the GPU)."""
    if a.spin_weight != b.spin_weight:
        raise ValueError("Spin weights must match in matrix_expectation_value")
    ell_min, ell_max = a.ell_min, a.ell_max
    clip = False
    if (a.ell_min != b.ell_min) or (a.ell_max != b.ell_max):
        if not allow_LM_differ:
            raise ValueError("ell_min and ell_max must match in matrix_expectation_value (use allow_LM_differ=True to override)")
        ell_min, ell_max = max(a.ell_min, b.ell_min), min(a.ell_max, b.ell_max)
        if ell_min >= ell_max + 1:
            raise ValueError("Intersection of (ell,m) modes is empty.  Assuming this is not desired.")
        clip = True
    t_clip = None
    if not np.array_equal(a.t, b.t):
        if not allow_times_differ:
            raise ValueError("Time samples must match in matrix_expectation_value (use allow_times_differ=True to override)")
        t_clip = intersection(a.t, b.t)
    times, A, B = a.t, a, b
    if clip:
        A, B = A[:, ell_min : ell_max + 1], B[:, ell_min : ell_max + 1]
    if t_clip is not None:
        times = t_clip
        A, B = A.interpolate(t_clip), B.interpolate(t_clip)
    rows, columns, values = M(ell_min, ell_max)
    return (times, ops.sparse_expectation(A.data, B.data, [(rows, columns, values)])[:, 0])


def _hdot_of(hw, name):
    from .waveform_modes import WaveformModes

    if not isinstance(hw, WaveformModes):
        raise ValueError(f"{name} can only be calculated from a `WaveformModes` object; this object is of type `{type(hw)}`.")
    if hw.dataType == hdottype:
        return hw.data
    if hw.dataType == htype:
        return _device_derivative(hw)[1]
    raise ValueError(f"Input argument is expected to have data of type `h` or `hdot`; this waveform data has type `{hw.data_type_string}`")


def _device_derivative(h):
    """(h.data, d/dt h.data) as device tensors: the modes travel to the device once and the derivative stays there
    (h.data_dot would bring it back to the host only for the flux kernels to send it again)."""
    h_d = ops.to_device(h.data, np.complex128)
    return h_d, ops.spline_calculus(ops.to_device(np.asarray(h.t, dtype=float), np.float64), h_d, "derivative", 1)


def energy_flux(h):
    """dE/dt = sum |hdot|^2 / 16 pi (scri/flux.py:182-210, Ruiz et al. 2008 eq. 2.8)"""
    hdot = _hdot_of(h, "Energy flux")
    return _energy_flux_of(hdot)


def _energy_flux_of(hdot):
    """`hdot` [N, n] as a host array or a device tensor; the result is a host array either way."""
    return ops.to_host(ops.norm(hdot)) / (16.0 * np.pi)


def momentum_flux(h):
    """dp/dt (scri/flux.py:303-342, Ruiz et al. 2008 eq. 2.11)"""
    hdot = _hdot_of(h, "Momentum flux")
    return _momentum_flux_of(hdot, h.ell_min, h.ell_max)


def _momentum_flux_of(hdot, lmin, lmax):
    ev = ops.to_host(ops.sparse_expectation(hdot, hdot, [p_plus(lmin, lmax, s=-2), p_minus(lmin, lmax, s=-2), p_z(lmin, lmax, s=-2)]))
    pdot = np.zeros((hdot.shape[0], 3), dtype=float)
    pdot[:, 0] = 0.5 * (ev[:, 0].real + ev[:, 1].real)
    pdot[:, 1] = 0.5 * (ev[:, 0].imag - ev[:, 1].imag)
    pdot[:, 2] = ev[:, 2].real
    pdot /= 16.0 * np.pi
    return pdot


def angular_momentum_flux(h, hdot=None):
    """dJ/dt (scri/flux.py:394-441, Ruiz et al. 2008 eq. 2.24)"""
    from .waveform_modes import WaveformModes

    if not isinstance(h, WaveformModes):
        raise ValueError(f"Angular momentum flux can only be calculated from a `WaveformModes` object; `h` is of type `{type(h)}`.")
    if (hdot is not None) and (not isinstance(hdot, WaveformModes)):
        raise ValueError(f"Angular momentum flux can only be calculated from a `WaveformModes` object; `hdot` is of type `{type(hdot)}`.")
    if h.dataType != htype:
        raise ValueError(f"Input argument `h` is expected to have data of type `h`; this `h` waveform data has type `{h.data_type_string}`")
    if hdot is None:
        h_data, hdot_data = _device_derivative(h)
    elif hdot.dataType != hdottype:
        raise ValueError(f"Input argument `hdot` is expected to have data of type `hdot`; this `hdot` waveform data has type `{hdot.data_type_string}`")
    else:
        h_data, hdot_data = ops.to_device(h.data, np.complex128), ops.to_device(hdot.data, np.complex128)
    return _angular_momentum_flux_of(h_data, hdot_data, h.ell_min, h.ell_max)


def _angular_momentum_flux_of(h_data, hdot_data, lmin, lmax):
    ev = ops.to_host(ops.sparse_expectation(hdot_data, h_data, [j_plus(lmin, lmax), j_minus(lmin, lmax), j_z(lmin, lmax)]))
    jdot = np.zeros((h_data.shape[0], 3), dtype=float)
    jdot[:, 0] = 0.5 * (ev[:, 0].real + ev[:, 1].real)
    jdot[:, 1] = 0.5 * (ev[:, 0].imag - ev[:, 1].imag)
    jdot[:, 2] = ev[:, 2].real
    jdot /= -16.0 * np.pi
    return jdot


def _ladder(operations, s, ell, eth_convention="NP"):
    from .waveform_modes import WaveformModes

    return WaveformModes.ladder_factor(None, operations, s, ell, eth_convention=eth_convention)


def _dressed(M, ell_min, ell_max, row_op, col_op, s, scale=1.0):
    """The matrix of <op_row a| M |op_col b> acting on the undressed modes: eth / ethbar only multiply each (l, m) mode
    by a real ladder factor, so  <f a|M|g b> = sum conj(a_r) (f_r M_rc g_c) b_c  and the 27 expectation values of the
    boost flux need no copies of eth h, ethbar h, ... - only matrices."""
    rows, cols, vals = M
    ells = np.concatenate([np.full(2 * ell + 1, ell) for ell in range(ell_min, ell_max + 1)])
    fr = np.array([_ladder(row_op, s, int(l)) if row_op else 1.0 for l in range(ell_min, ell_max + 1)])
    fc = np.array([_ladder(col_op, s, int(l)) if col_op else 1.0 for l in range(ell_min, ell_max + 1)])
    return rows, cols, scale * fr[ells[rows] - ell_min] * vals * fc[ells[cols] - ell_min]


@functools.lru_cache(maxsize=None)
def eth_chi_z(ell_min, ell_max, s=-2):
    """<eth N|eth chi|h>, z direction (scri/flux.py:486-505)"""
    out = []
    for ell in range(ell_min, ell_max + 1):
        for ellp in range(max(ell_min, ell - 1), min(ell_max, ell + 1) + 1):
            cg2 = clebsch_gordan(ell, -s, 1, -1, ellp, -1 - s)
            pre = math.sqrt((2.0 * ell + 1.0) / (2.0 * ellp + 1.0))
            for m in range(-ell, ell + 1):
                if abs(m) > ellp:
                    continue
                out.append((ellp, m, ell, m, pre * math.sqrt(2) * clebsch_gordan(ell, m, 1, 0, ellp, m) * cg2))
    return _as_matrix(out, ell_min)


@functools.lru_cache(maxsize=None)
def ethbar_chi_z(ell_min, ell_max, s=-2):
    """<h|ethbar chi|eth N>, z direction (scri/flux.py:507-526)"""
    out = []
    for ell in range(ell_min, ell_max + 1):
        for ellp in range(max(ell_min, ell - 1), min(ell_max, ell + 1) + 1):
            cg2 = clebsch_gordan(ell, -s - 1, 1, 1, ellp, -s)
            pre = math.sqrt((2.0 * ell + 1.0) / (2.0 * ellp + 1.0))
            for m in range(-ell, ell + 1):
                if abs(m) > ellp:
                    continue
                out.append((ellp, m, ell, m, pre * math.sqrt(2) * clebsch_gordan(ell, m, 1, 0, ellp, m) * cg2))
    return _as_matrix(out, ell_min)


@functools.lru_cache(maxsize=None)
def _chi_plusminus(ell_min, ell_max, sign, s, bar):
    if sign not in (1, -1):
        raise ValueError("sign must be either 1 or -1 in eth_chi_plusminus")
    prefac = -1.0 * sign * math.sqrt(8.0 * math.pi / 3.0)
    out = []
    for ell in range(ell_min, ell_max + 1):
        for ellp in range(max(ell_min, ell - 1), min(ell_max, ell + 1) + 1):
            for m in range(-ell, ell + 1):
                mp = m + sign
                if abs(mp) > ellp:
                    continue
                cg1 = math.sqrt(2) * clebsch_gordan(1, sign, ell, m, ellp, mp)
                cg2 = clebsch_gordan(1, 1, ell, -1 - s, ellp, -s) if bar else clebsch_gordan(1, -1, ell, -s, ellp, -1 - s)
                el = math.sqrt(3.0 * (2.0 * ell + 1.0) / (4.0 * math.pi * (2.0 * ellp + 1))) * cg1 * cg2
                out.append((ellp, mp, ell, m, prefac * el))
    return _as_matrix(out, ell_min)


def eth_chi_plusminus(ell_min, ell_max, sign, s=-2):
    """<eth N|eth chi|h>, m = +-1 (scri/flux.py:528-568)"""
    return _chi_plusminus(ell_min, ell_max, sign, s, False)


def ethbar_chi_plusminus(ell_min, ell_max, sign, s=-2):
    """<h|ethbar chi|eth N>, m = +-1 (scri/flux.py:575-615)"""
    return _chi_plusminus(ell_min, ell_max, sign, s, True)


def boost_flux(h, hdot=None):
    """Boost flux, Flanagan & Nichols (2016) eq. C.1 (scri/flux.py:444-747).

    The reference evaluates 27 `matrix_expectation_value`s on six copies of the data (h, eth h, ethbar h and the same
    for hdot).  Here the ladder factors are folded into the sparse matrices and the terms are grouped by operand pair,
    so the modes are read in three fused passes: <hdot|.|h> (15 matrices), <h|.|hdot> (12) and <hdot|.|hdot> (3).
    """
    from .waveform_modes import WaveformModes

    if not isinstance(h, WaveformModes):
        raise ValueError(f"Boost fluxes can only be calculated from a `WaveformModes` object; `h` is of type `{type(h)}`.")
    if (hdot is not None) and (not isinstance(hdot, WaveformModes)):
        raise ValueError(f"Boost fluxes can only be calculated from a `WaveformModes` object; `hdot` is of type `{type(hdot)}`.")
    if h.dataType != htype:
        raise ValueError(f"Input argument `h` is expected to have data of type `h`; this `h` waveform data has type `{h.data_type_string}`")
    if hdot is None:
        h_data, hdot_data = _device_derivative(h)
    elif hdot.dataType != hdottype:
        raise ValueError(f"Input argument `hdot` is expected to have data of type `hdot`; this `hdot` waveform data has type `{h.data_type_string}`")
    else:
        h_data, hdot_data = ops.to_device(h.data, np.complex128), ops.to_device(hdot.data, np.complex128)
    return _boost_flux_of(h_data, hdot_data, np.asarray(h.t), h.ell_min, h.ell_max)


@functools.lru_cache(maxsize=8)
def _boost_matrices(lo, hi):
    """The three groups of sparse matrices of the boost flux, ladder factors folded in (built once per ell range)."""
    s = -2
    comps = []
    for P, EC, EBC in (
        (functools.partial(p_plusminus, lo, hi, +1), eth_chi_plusminus(lo, hi, +1, s), ethbar_chi_plusminus(lo, hi, +1, s)),
        (functools.partial(p_plusminus, lo, hi, -1), eth_chi_plusminus(lo, hi, -1, s), ethbar_chi_plusminus(lo, hi, -1, s)),
        (functools.partial(p_z, lo, hi), eth_chi_z(lo, hi, s), ethbar_chi_z(lo, hi, s)),
    ):
        comps.append(dict(
            nh=[_dressed(P(s=-3), lo, hi, "-", "-", s), _dressed(P(s=-1), lo, hi, "+", "+", s), P(s=-2), _dressed(EC, lo, hi, "+", "", s)],
            hn=[_dressed(P(s=-3), lo, hi, "-", "-", s), _dressed(P(s=-1), lo, hi, "+", "+", s), P(s=-2), _dressed(EBC, lo, hi, "", "+", s)],
            nn=[P(s=-2)],
        ))
    return tuple([m for c in comps for m in c[k]] for k in ("nh", "hn", "nn"))


def _boost_flux_of(h_data, hdot_data, t, lo, hi):
    m_nh, m_hn, m_nn = _boost_matrices(lo, hi)
    ev_nh = ops.to_host(ops.sparse_expectation(hdot_data, h_data, m_nh))      # <hdot| . |h>
    ev_hn = ops.to_host(ops.sparse_expectation(h_data, hdot_data, m_hn))      # <h| . |hdot>
    ev_nn = ops.to_host(ops.sparse_expectation(hdot_data, hdot_data, m_nn))   # <hdot| . |hdot>
    total = []
    for i in range(3):
        nh, hn = ev_nh[:, 4 * i : 4 * i + 4], ev_hn[:, 4 * i : 4 * i + 4]
        first = (1 / 8) * (nh[:, 0] - nh[:, 1] + 6 * nh[:, 2] + hn[:, 0] - hn[:, 1] + 6 * hn[:, 2])
        total.append(first - (1 / 2) * t * ev_nn[:, i] - (1 / 4) * (nh[:, 3] - hn[:, 3]))
    out = np.zeros((h_data.shape[0], 3), dtype=float)
    out[:, 0] = 0.5 * (total[0] + total[1]).real
    out[:, 1] = 0.5 * (total[0] - total[1]).imag
    out[:, 2] = total[2].real
    out /= -32 * np.pi
    return out


def poincare_fluxes(h, hdot=None):
    """(energy, momentum, angular-momentum, boost) fluxes with one time derivative (scri/flux.py:750-797).  The modes go
    to the device once: the derivative is taken there and the four fluxes read the two device arrays (the reference makes
    a full copy of the waveform to hold hdot; nothing of the kind is needed here)."""
    from .waveform_modes import WaveformModes

    if not isinstance(h, WaveformModes):
        raise ValueError(f"Poincare fluxes can only be calculated from a `WaveformModes` object; `h` is of type `{type(h)}`.")
    if h.dataType != htype:
        raise ValueError(f"Input argument `h` is expected to have data of type `h`; this `h` waveform data has type `{h.data_type_string}`")
    if hdot is None:
        h_d, hdot_d = _device_derivative(h)
    else:
        h_d = ops.to_device(h.data, np.complex128)
        if not isinstance(hdot, WaveformModes):
            raise ValueError(f"Poincare fluxes can only be calculated from a `WaveformModes` object; `hdot` is of type `{type(hdot)}`.")
        if hdot.dataType != hdottype:
            raise ValueError(f"Input argument `hdot` is expected to have data of type `hdot`; this `hdot` waveform data has type `{hdot.data_type_string}`")
        hdot_d = ops.to_device(hdot.data, np.complex128)
    lo, hi = h.ell_min, h.ell_max
    return (_energy_flux_of(hdot_d), _momentum_flux_of(hdot_d, lo, hi), _angular_momentum_flux_of(h_d, hdot_d, lo, hi),
            _boost_flux_of(h_d, hdot_d, np.asarray(h.t), lo, hi))
