"""numpy emulation of csrc/product.cu's three stages, driven by the SAME host tables (scri_b200._product): checks the
fragment layouts, the mode permutation and the convolution/alias bookkeeping on a CPU.  Test infrastructure only."""
import numpy as np

from scri_b200 import _product


def emulate(tb, a1, a2):
    cfg = [int(x) for x in tb.cfg]
    ell1, ell2, L_out, n_phi, n_chunks, qmax, szA, offF1, offF2, smem, nwarps, max_ks, GM, maxt, _ = cfg
    N = a1.shape[0]
    n_out = (L_out + 1) ** 2
    out = np.zeros((N, n_out), dtype=complex)
    for t0 in range(0, N, 4):
        nt = min(4, N - t0)
        sm = np.zeros(smem)
        for t in range(nt):
            for a, perm in ((a1, tb.perm1), (a2, tb.perm2)):
                sm[perm + 2 * t] = a[t0 + t].real
                sm[perm + 2 * t + 1] = a[t0 + t].imag
        acc = np.zeros((len(tb.tiles), 8, 8))
        for c in range(n_chunks):
            ctl = tb.ctl.view(np.uint32)
            woff = ctl[: nwarps + 1]
            for wp in range(nwarps):
                C = np.zeros((8, 8))
                for u in ctl[woff[wp] : woff[wp + 1]]:
                    g = int(u) & 0xFFFF
                    A = tb.lamfrag[c, 32 * g : 32 * g + 32].reshape(8, 4)
                    B = sm[32 * g : 32 * g + 32].reshape(4, 8)
                    C += A @ B
                    if int(u) >> 31:
                        w = offF1 + 64 * ((int(u) >> 16) & 0x7FFF)
                        sm[w : w + 64] = C.reshape(64)
                        C = np.zeros((8, 8))
            F1 = sm[offF1 : offF1 + 64 * (2 * ell1 + 1)].reshape(-1, 32, 2)
            F1 = F1[..., 0] + 1j * F1[..., 1]
            F2 = sm[offF2 : offF2 + 64 * (2 * ell2 + 1)].reshape(-1, 32, 2)
            F2 = F2[..., 0] + 1j * F2[..., 1]
            P = np.zeros((2 * L_out + 1, 32), dtype=complex)
            for Mi in range(2 * L_out + 1):
                M = Mi - L_out
                for q in range(-qmax, qmax + 1):
                    Me = M + q * n_phi
                    for m1 in range(max(-ell1, Me - ell2), min(ell1, Me + ell2) + 1):
                        P[Mi] += F1[m1 + ell1] * F2[Me - m1 + ell2]
            sm[offF1 : offF1 + 64 * (2 * L_out + 1)] = np.stack([P.real, P.imag], axis=-1).reshape(-1)
            for ti, (Mi, l0) in enumerate(tb.tiles):
                for k in range(2):
                    A = tb.wtfrag[c, ti * 64 : ti * 64 + 64].reshape(8, 4, 2)[:, :, k]
                    B = sm[offF1 + Mi * 64 + 32 * k : offF1 + Mi * 64 + 32 * k + 32].reshape(4, 8)
                    acc[ti] += A @ B
        for ti, (Mi, l0) in enumerate(tb.tiles):
            M = Mi - L_out
            for r in range(8):
                l = l0 + r
                if l <= L_out:
                    for t in range(nt):
                        out[t0 + t, l * (l + 1) + M] = acc[ti, r, 2 * t] + 1j * acc[ti, r, 2 * t + 1]
    return out
