"""numpy emulation of csrc/product.cu's three stages, driven by the SAME host tables (scri_b200._product): checks the
fragment layouts, the mode permutation and the convolution/alias bookkeeping on a CPU.  Test infrastructure only."""
import numpy as np

from scri_b200 import _product


def emulate_cluster(tb, a1, a2):
    """The cluster variant: two CTA ranks, each with its own mode tile, control streams, M groups and output tiles;
    stage A of either rank lands in the F buffers of both (modelled as one shared buffer)."""
    cfg = [int(x) for x in tb.cfg]
    ell1, ell2, L_out, n_phi, n_chunks, qmax, szA, bufStride, offF2rel, smem = cfg[:10]
    GM, maxt = cfg[12], cfg[13]
    gbase = [0, cfg[16]]
    g0, gcnt = [cfg[17], cfg[19]], [cfg[18], cfg[20]]
    N = a1.shape[0]
    n_out = (L_out + 1) ** 2
    n_mout = 2 * L_out + 1
    out = np.zeros((N, n_out), dtype=complex)
    ctl = tb.ctl.view(np.uint32).reshape(2, tb.n_ctl)
    tiles = tb.tiles.reshape(2, -1, 2)
    tiles_r = tiles.shape[1]
    for t0 in range(0, N, 4):
        nt = min(4, N - t0)
        sA = [np.zeros(smem), np.zeros(smem)]           # per-rank shared memory (only the mode tile part differs)
        for r, (a, perm) in enumerate(((a1, tb.perm1), (a2, tb.perm2))):
            for t in range(nt):
                sA[r][perm + 2 * t] = a[t0 + t].real
                sA[r][perm + 2 * t + 1] = a[t0 + t].imag
        acc = np.zeros((2, tiles_r, 8, 8))
        for c in range(n_chunks):
            buf = np.zeros(bufStride)
            for r in range(2):
                woff = ctl[r, :9]
                assert sorted(int(x) for x in ctl[r, 9:17] if x != 0xFFFFFFFF) == list(range(g0[r], g0[r] + gcnt[r]))
                for wp in range(8):
                    C = np.zeros((3, 8, 8))                  # three chains: stream position i feeds chain i % 3
                    for i, u in enumerate(ctl[r, woff[wp] : woff[wp + 1]]):
                        g = int(u) & 0xFFFF
                        A = tb.lamfrag[c, 32 * g : 32 * g + 32].reshape(8, 4)
                        B = sA[r][32 * (g - gbase[r]) : 32 * (g - gbase[r]) + 32].reshape(4, 8)
                        C[i % 3] += A @ B
                        if int(u) >> 31:
                            w = 64 * ((int(u) >> 16) & 0x7FFF)
                            buf[w : w + 64] = C[i % 3].reshape(64)
                            C[i % 3] = 0.0
            F1 = buf[: 64 * (2 * ell1 + 1)].reshape(-1, 32, 2)
            F1 = F1[..., 0] + 1j * F1[..., 1]
            F2 = buf[offF2rel : offF2rel + 64 * (2 * ell2 + 1)].reshape(-1, 32, 2)
            F2 = F2[..., 0] + 1j * F2[..., 1]
            P = np.zeros((n_mout, 32), dtype=complex)
            done = np.zeros(n_mout, dtype=bool)
            for r in range(2):
                for gq in range(g0[r], g0[r] + gcnt[r]):
                    for Mi in range(GM * gq, min(GM * gq + GM, n_mout)):
                        M = Mi - L_out
                        assert not done[Mi]
                        done[Mi] = True
                        for q in range(-qmax, qmax + 1):
                            Me = M + q * n_phi
                            for m1 in range(max(-ell1, Me - ell2), min(ell1, Me + ell2) + 1):
                                P[Mi] += F1[m1 + ell1] * F2[Me - m1 + ell2]
            assert done.all()
            Pbuf = np.stack([P.real, P.imag], axis=-1).reshape(-1)
            for r in range(2):
                for ti, (Mi, l0) in enumerate(tiles[r]):
                    if l0 > L_out:
                        continue
                    assert g0[r] <= Mi // GM < g0[r] + gcnt[r]
                    base = ((c * 2 + r) * tiles_r + ti) * 64
                    for k in range(2):
                        A = tb.wtfrag.reshape(-1)[base : base + 64].reshape(8, 4, 2)[:, :, k]
                        B = Pbuf[Mi * 64 + 32 * k : Mi * 64 + 32 * k + 32].reshape(4, 8)
                        acc[r, ti] += A @ B
        for r in range(2):
            for ti, (Mi, l0) in enumerate(tiles[r]):
                M = Mi - L_out
                for rr in range(8):
                    l = l0 + rr
                    if l <= L_out:
                        for t in range(nt):
                            out[t0 + t, l * (l + 1) + M] = acc[r, ti, rr, 2 * t] + 1j * acc[r, ti, rr, 2 * t + 1]
    return out


def emulate(tb, a1, a2):
    if getattr(tb, "cluster", False):
        return emulate_cluster(tb, a1, a2)
    cfg = [int(x) for x in tb.cfg]
    ell1, ell2, L_out, n_phi, n_chunks, qmax, szA, offF1, offF2, smem, nwarps, max_ks, GM, maxt = cfg[:14]
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
            assert sorted(int(x) for x in ctl[nwarps + 1 : 2 * nwarps + 1] if x != 0xFFFFFFFF) == list(range(-(-(2 * L_out + 1) // GM)))
            for wp in range(nwarps):
                C = np.zeros((3, 8, 8))                      # three chains: stream position i feeds chain i % 3
                for i, u in enumerate(ctl[woff[wp] : woff[wp + 1]]):
                    g = int(u) & 0xFFFF
                    A = tb.lamfrag[c, 32 * g : 32 * g + 32].reshape(8, 4)
                    B = sm[32 * g : 32 * g + 32].reshape(4, 8)
                    C[i % 3] += A @ B
                    if int(u) >> 31:
                        w = offF1 + 64 * ((int(u) >> 16) & 0x7FFF)
                        sm[w : w + 64] = C[i % 3].reshape(64)
                        C[i % 3] = 0.0
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
