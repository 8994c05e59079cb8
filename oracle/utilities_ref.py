"""CPU restatement of the integer stages of scri's RPXMB codec: scri/utilities.py:194-407.

TEST INFRASTRUCTURE (see oracle/__init__.py) - never imported by scri_b200/.  Plain numpy, written from the documented
behaviour rather than from the numba loops: the XOR of successive time steps and its inverse (:194-232), Fletcher-32 over 16-bit
words with modulus 65535 (:235-274), and the bit-level multishuffle (:277-407).  Pinned by the reference's own known answer: with
widths (8,) * (bits / 8) the multishuffle must equal HDF5's byte shuffle (tests/test_utilities.py:31-52), whose layout - byte 0 of
every element, then byte 1 of every element, ... - is the published filter definition; and by reversibility (:20-28).
"""
import numpy as np


def xor_timeseries(c):
    """out[i] = c[i] ^ c[i-1] on the 64-bit patterns, first time step unchanged (utilities.py:194-215); returns a new array."""
    u = np.ascontiguousarray(c).view(np.uint64)
    out = u.copy()
    out[1:] = u[1:] ^ u[:-1]
    return out.view(c.dtype).reshape(np.shape(c))


def xor_timeseries_reverse(c):
    """Prefix XOR along the first axis: undoes xor_timeseries bit for bit (utilities.py:218-232)."""
    u = np.ascontiguousarray(c).view(np.uint64)
    return np.bitwise_xor.accumulate(u, axis=0).view(c.dtype).reshape(np.shape(c))


def fletcher32(data):
    """16-bit words, 32-bit sums reduced modulo 65535 every 360 words, c1 << 16 | c0 (utilities.py:235-274)."""
    d = np.ascontiguousarray(data).reshape(-1).view(np.uint16).astype(np.uint64)
    c0 = c1 = 0
    for j in range(0, d.size, 360):
        blk = d[j : j + 360]
        run = c0 + np.cumsum(blk)
        c1 = (c1 + int(run.sum())) % 65535
        c0 = int(run[-1]) % 65535
    return np.uint32((c1 << 16) | c0)


def multishuffle(shuffle_widths, forward=True):
    """Shuffle function for the given bit widths (highest significance first), as utilities.py:277-407 builds it: the output is
    the little-endian bit stream holding the LOWEST piece of every element, then the next piece of every element, ...;
    forward=False returns the inverse."""
    bit_width = int(np.sum(shuffle_widths))
    if bit_width not in (8, 16, 32, 64):
        raise ValueError(f"Total bit width must be one of [8, 16, 32, 64], not {bit_width}")
    dtype = np.dtype(f"u{bit_width // 8}")
    widths = list(reversed([int(w) for w in shuffle_widths]))
    shifts = np.concatenate([[0], np.cumsum(widths)])[:-1]

    def bits_of(a):          # [n, bit_width] little-endian bits
        return ((a[:, None].astype(np.uint64) >> np.arange(bit_width, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8)

    def from_bits(bits):
        weights = (np.uint64(1) << np.arange(bit_width, dtype=np.uint64))
        return (bits.astype(np.uint64) * weights[None, :]).sum(axis=1, dtype=np.uint64).astype(dtype)

    def shuffle(a):
        a = np.ascontiguousarray(a).view(dtype)
        if a.ndim != 1:
            raise ValueError("This function only accepts flat arrays.")
        ab = bits_of(a)
        stream = np.concatenate([ab[:, s : s + w].reshape(-1) for s, w in zip(shifts, widths)])
        return from_bits(stream.reshape(a.size, bit_width))

    def unshuffle(b):
        b = np.ascontiguousarray(b).view(dtype)
        if b.ndim != 1:
            raise ValueError("This function only accepts flat arrays.")
        stream = bits_of(b).reshape(-1)
        ab = np.zeros((b.size, bit_width), dtype=np.uint8)
        pos = 0
        for s, w in zip(shifts, widths):
            ab[:, s : s + w] = stream[pos : pos + b.size * w].reshape(b.size, w)
            pos += b.size * w
        return from_bits(ab)

    return shuffle if forward else unshuffle
