"""Integer stages of the RPXMB waveform codec on the GPU: the drop-in for scri/utilities.py:194-407 (`xor_timeseries`,
`xor_timeseries_reverse`, `fletcher32`, `multishuffle`) as scri/SpEC/file_io/corotating_paired_xor.py and
rotating_paired_xor_multishuffle_bzip2.py call them, and `corotating_paired_xor_encode`, the numerical part of the former's
`save`.  Bit-exact; numpy arrays in, numpy arrays out.  No CPU fallback."""
import ctypes
import functools

import numpy as np

from . import _lib, ops


def _as_u64_rows(c):
    a = np.ascontiguousarray(c)
    if a.dtype.itemsize * int(np.prod(a.shape[1:], dtype=np.int64)) % 8:
        raise ValueError("xor_timeseries: every time step must be a whole number of 64-bit words")
    n_rows = a.shape[0]
    return a, n_rows, (a.size * a.dtype.itemsize // 8) // max(n_rows, 1)


def _xor(c, reverse):
    torch = _lib.require_cuda()
    lib = _lib.load()
    a, n_rows, n_cols = _as_u64_rows(c)
    if a.size == 0:
        return c
    d = ops.to_device(a.reshape(-1).view(np.uint8))
    out = torch.empty_like(d)
    need = lib.scrib200_xor_timeseries_workspace_bytes(n_rows, n_cols)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    _lib.check(lib.scrib200_xor_timeseries(_lib.ptr(d), _lib.ptr(out), n_rows, n_cols, int(reverse), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
               "xor_timeseries")
    res = ops.to_host(out).view(a.dtype).reshape(a.shape)
    if isinstance(c, np.ndarray) and c.flags.writeable and c.shape == res.shape and c.dtype == res.dtype:
        c[...] = res          # the reference works in place and returns its argument
        return c
    return res


def xor_timeseries(c):
    """XOR every time step (first axis) with its predecessor, in place; the first stays (scri/utilities.py:194-215)."""
    return _xor(c, False)


def xor_timeseries_reverse(c):
    """Undo `xor_timeseries` bit for bit (scri/utilities.py:218-232)."""
    return _xor(c, True)


def fletcher32(data):
    """Fletcher-32 checksum over the 16-bit words of `data` (scri/utilities.py:235-274): uint32 c1 << 16 | c0."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    a = np.ascontiguousarray(data).reshape(-1)
    if (a.size * a.dtype.itemsize) % 2:
        raise ValueError("fletcher32: the data must be viewable as 16-bit words")
    acc = torch.zeros(2, dtype=torch.int64, device="cuda")
    n16 = a.size * a.dtype.itemsize // 2
    if n16:
        d = ops.to_device(a.view(np.uint8))
        _lib.check(lib.scrib200_fletcher32(_lib.ptr(d), n16, _lib.ptr(acc), _lib.stream_ptr()), "fletcher32")
    c0, c1 = (int(x) % 65535 for x in acc.cpu().tolist())
    return np.uint32((c1 << 16) | c0)


@functools.lru_cache()
def multishuffle(shuffle_widths, forward=True):
    """Function that multishuffles (or, with forward=False, unshuffles) a flat array of 8 / 16 / 32 / 64-bit numbers: the bits of
    every element are cut into pieces of the given widths (highest significance first) and like pieces are stored together,
    starting with the lowest (scri/utilities.py:277-407)."""
    widths = tuple(int(w) for w in shuffle_widths)
    bit_width = int(np.sum(widths))
    if bit_width not in (8, 16, 32, 64):
        raise ValueError(f"Total bit width must be one of [8, 16, 32, 64], not {bit_width}")
    dtype = np.dtype(f"u{bit_width // 8}")
    w_arr = (ctypes.c_int * len(widths))(*widths)

    def run(a):
        torch = _lib.require_cuda()
        lib = _lib.load()
        a = np.ascontiguousarray(a).view(dtype)
        if a.ndim != 1:
            raise ValueError(
                "\nThis function only accepts flat arrays.  Make sure you flatten (using ravel, reshape, or flatten)\n"
                " in a way that keeps your data contiguous in the order you want."
            )
        if a.size == 0:
            return a.copy()
        d = ops.to_device(a.view(np.uint8))
        out = torch.empty_like(d)
        _lib.check(lib.scrib200_multishuffle(_lib.ptr(d), _lib.ptr(out), a.size, bit_width, w_arr, len(widths), int(forward), _lib.stream_ptr()),
                   "multishuffle")
        return ops.to_host(out).view(dtype)

    return run


def corotating_paired_xor_encode(w, L2norm_fractional_tolerance=1e-10, log_frame=None):
    """The numerical part of scri/SpEC/file_io/corotating_paired_xor.py:save (lines 46-91, 121-126) on the GPU path: the
    waveform goes to the corotating frame if it is inertial (tolerance 1e-10, z aligned over (0.1, 0.95), log of the frame
    on its lattice), to conjugate pairs, is truncated to `L2norm_fractional_tolerance` of its norm, -0.0 becomes 0.0, and
    successive instants are XORed.  Returns (w, streams, checksums): `w` the encoded copy, `streams` = {"time", "modes",
    "log_frame"} as the uint64 arrays the reference writes to HDF5, `checksums` their Fletcher-32 values (the "validation"
    block of the JSON file).  The HDF5 / JSON container itself stays with the reference."""
    from . import _quaternion as Q
    from .constants import Corotating, Inertial

    if L2norm_fractional_tolerance == 0.0:
        log_frame = Q.qlog(w.frame)[:, 1:]
    else:
        w = w.copy()
        if log_frame is not None:
            log_frame = np.array(log_frame, dtype=float)
        if w.frameType == Inertial:
            w, log_frame = w.to_corotating_frame(tolerance=1e-10, z_alignment_region=(0.1, 0.95), truncate_log_frame=True)
            log_frame = log_frame[:, 1:]
        if w.frameType != Corotating:
            raise ValueError(f"Frame type of input waveform must be 'Corotating' or 'Inertial'; it is {w.frame_type_string}")
        w.convert_to_conjugate_pairs()
        w.truncate(tol=L2norm_fractional_tolerance)
        if log_frame is None:
            log_frame = Q.qlog(w.frame)[:, 1:]
            power_of_2 = 2 ** (-np.floor(np.log2(L2norm_fractional_tolerance / 10))).astype("int")
            log_frame = np.round(log_frame * power_of_2) / power_of_2
        w.t = w.t + 0.0
        w.data = w.data + 0.0
        log_frame = np.ascontiguousarray(log_frame + 0.0)
        w.t = xor_timeseries(np.ascontiguousarray(w.t))
        w.data = xor_timeseries(np.ascontiguousarray(w.data))
        log_frame = xor_timeseries(log_frame)
    streams = {"time": w.t.view(np.uint64), "modes": w.data.view(np.uint64), "log_frame": np.ascontiguousarray(log_frame).view(np.uint64)}
    checksums = {k: int(fletcher32(v)) for k, v in streams.items()}
    return w, streams, checksums
