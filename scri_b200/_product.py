"""Host tables of the fused, separable mode product (K9, csrc/product.cu).

`ModesTimeSeries.grid_multiply` (scri/modes_time_series.py:142-202) evaluates both factors on spinsfast's regular
(theta, phi) grid, multiplies pointwise and analyses the product.  On that grid sY_lm(theta_j, phi_k) =
lambda_lm(theta_j) e^{i m phi_k}, so the whole chain is, ring by ring, a Wigner-d contraction over l, a convolution
over m and a theta quadrature; the kernel needs the lambda and quadrature tables laid out as DMMA fragments and the
mode permutation that makes the l of one m contiguous in shared memory.  Pure layout work: no data-path arithmetic.
"""
from functools import lru_cache

import numpy as np

from . import _sf

T = 4           # time steps per CTA pass (PRODUCT_T)
# kernel shapes instantiated in csrc/product.cu: (consecutive M per convolution warp, warps, output tiles per warp, fragment ring depth)
SHAPES = ((5, 16, 11, 9), (9, 8, 21, 18))
MAX_SMEM = 227 * 1024


class ProductTables:
    """Numpy tables + the configuration vector of scrib200_modes_product; `fits` is False when one CTA cannot hold the
    problem (callers then use the dense synthesis / analysis kernels)."""


def _lambda(s, ell_max, n_theta):
    """lambda_lm(theta_j) = sY_lm(theta_j, 0) on spinsfast's rings theta_j = pi j / (n_theta - 1): [n_theta, (ell_max+1)^2]."""
    theta = np.pi * np.arange(n_theta) / (n_theta - 1)
    R = np.stack([np.cos(theta / 2), np.zeros_like(theta), np.sin(theta / 2), np.zeros_like(theta)], axis=-1)
    return np.ascontiguousarray(_sf.SWSH_grid(R, s, ell_max).real)


def _field_layout(ell_min, ell_max):
    """m-major layout of one factor: for every m the l = max(|m|, ell_min) .. ell_max are contiguous and padded to a
    multiple of 4 (one DMMA k-step).  Returns (pos[m + ell_max], ksteps[m + ell_max], padded size, perm[idx])."""
    ms = np.arange(-ell_max, ell_max + 1)
    l_lo = np.maximum(np.abs(ms), ell_min)
    ks = (ell_max - l_lo + 1 + 3) // 4
    pos = np.concatenate([[0], np.cumsum(4 * ks)])
    perm = np.empty((ell_max + 1) ** 2 - ell_min**2, dtype=np.int64)
    for l in range(ell_min, ell_max + 1):
        m = np.arange(-l, l + 1)
        perm[l * (l + 1) + m - ell_min**2] = pos[m + ell_max] + (l - l_lo[m + ell_max])
    return pos[:-1], ks, int(pos[-1]), perm, l_lo


def _build_streams(tasks, nwarps, DA, nop):
    """Stage A control streams.  Every warp runs three DMMA chains side by side: the k-steps of three (field, m) tasks are
    interleaved (stream position i belongs to chain i % 3), so each accumulator holds exactly one task and a finished task is
    one predicated store.  Tasks go longest first to the shortest of the 3 * nwarps queues; streams are padded with no-op
    steps (an all-zero fragment) to a multiple of the kernel's fragment ring."""
    queues = [[] for _ in range(3 * nwarps)]
    for g0, ks, entry in sorted(tasks, key=lambda x: -x[1]):
        q = min(queues, key=len)
        q.extend((g0 + k) | (entry << 16) | ((1 << 31) if k == ks - 1 else 0) for k in range(ks))
    streams = []
    for w in range(nwarps):
        qs = queues[w::nwarps]
        st = [qs[j][i] if i < len(qs[j]) else nop for i in range(max(len(q) for q in qs)) for j in range(3)]
        st.extend([nop] * (-len(st) % DA))
        streams.append(st)
    return streams


def _assign_groups(groups, n_slots, GM, L_out, ell1, ell2, n_phi, qmax):
    """Convolution groups (GM consecutive M each) -> warp slots.  A group's work is its number of (m1, m2) blocks; warp
    slot w runs on SM sub-partition w % 4, so groups go heaviest first to the least loaded sub-partition with a free slot.
    Returns the group of every slot (0xffffffff = idle)."""
    def work(g):
        M0 = -L_out + GM * g
        tot = 0
        for q in range(-qmax, qmax + 1):
            Me = M0 + q * n_phi
            lo, hi = max(-ell1, Me - ell2), min(ell1, Me + GM - 1 + ell2)
            tot += max(0, hi - lo + 1)
        return tot

    free = {sp: [w for w in range(n_slots) if w % 4 == sp] for sp in range(4)}
    load = {sp: 0 for sp in range(4)}
    out = [0xFFFFFFFF] * n_slots
    for g in sorted(groups, key=lambda g: -work(g)):
        sp = min((sp for sp in range(4) if free[sp]), key=lambda sp: load[sp])
        out[free[sp].pop(0)] = g
        load[sp] += work(g)
    return out


@lru_cache(maxsize=16)
def product_tables(s1, ell1_min, ell1_max, s2, ell2_min, ell2_max, n_theta, n_phi, L_out, shape=None):
    if shape == 2:
        return _cluster_tables(s1, ell1_min, ell1_max, s2, ell2_min, ell2_max, n_theta, n_phi, L_out)
    tb = ProductTables()
    tb.cluster = False
    n_chunks = (n_theta + 7) // 8
    n_rings = 8 * n_chunks
    n_mout = 2 * L_out + 1
    fields = []
    base = 0
    for s, lmin, lmax in ((s1, ell1_min, ell1_max), (s2, ell2_min, ell2_max)):
        pos, ks, PA, perm, l_lo = _field_layout(lmin, lmax)
        fields.append(dict(s=s, lmin=lmin, lmax=lmax, pos=pos, ks=ks, PA=PA, perm=perm, l_lo=l_lo, base=base))
        base += PA + 4                                      # one all-zero k-step after each factor: the target of no-op steps
    PA_total = base
    szA = 8 * PA_total                                     # doubles: [pos][t = 4][re, im]
    offF1 = szA
    nF1 = max(2 * ell1_max + 1, n_mout)                     # the product P_M overwrites F1
    gm_max = max(sh[0] for sh in SHAPES)
    offF2 = offF1 + 64 * nF1 + 64 * (gm_max + 2)           # zero guards around F2: GM + 2 entries below, GM - 1 above
    smem_doubles = offF2 + 64 * (2 * ell2_max + 1) + 64 * (gm_max - 1)

    # output tiles: 8 consecutive l of one M
    tiles = [(M + L_out, l0) for M in range(-L_out, L_out + 1) for l0 in range(abs(M), L_out + 1, 8)]
    n_steps = PA_total // 4
    chosen = None
    for sh in SHAPES if shape is None else (SHAPES[shape],):
        gm, nw, maxt, _ = sh
        if -(-n_mout // gm) <= nw and len(tiles) <= nw * maxt:
            chosen = sh
            break
    tb.fits = chosen is not None and 8 * smem_doubles + 4 * n_steps <= MAX_SMEM and n_steps < 65535 and n_theta >= 2
    tb.smem_bytes = 8 * smem_doubles + 4 * n_steps
    if tb.fits:
        GM, nwarps, maxt, DA = chosen
        tiles += [(0, L_out + 1)] * (nwarps * maxt - len(tiles))     # empty tiles: every warp walks `maxt` of them
    n_tiles = len(tiles)
    if not tb.fits:
        return tb

    # stage A: lambda fragments.  lam_pad[ring, base + pos] in the m-major padded order, then per (chunk, k-step) the
    # 8 rings x 4 l block in lane order (row = lane/4 = ring, k = lane%4 = l)
    lam_pad = np.zeros((n_rings, PA_total))
    tasks = []
    for fi, f in enumerate(fields):
        lam = _lambda(f["s"], f["lmax"], n_theta)           # [n_theta, (lmax+1)^2], zero for l < |s|
        n = f["perm"].shape[0]
        lam_pad[:n_theta, f["base"] + f["perm"]] = lam[:, f["lmin"] ** 2 : f["lmin"] ** 2 + n]
        entry0 = 0 if fi == 0 else (offF2 - offF1) // 64
        for mi in range(2 * f["lmax"] + 1):
            tasks.append(((f["base"] + int(f["pos"][mi])) // 4, int(f["ks"][mi]), entry0 + mi))   # (first k-step, k-steps, F entry)
    nop = n_steps - 1                                        # the zero k-step after the second factor: zero fragment, zero modes
    streams = _build_streams(tasks, nwarps, DA, nop)
    qmax = (ell1_max + ell2_max + L_out) // n_phi
    grp = _assign_groups(range(-(-n_mout // GM)), nwarps, GM, L_out, ell1_max, ell2_max, n_phi, qmax)
    woff = (2 * nwarps + 1) + np.concatenate([[0], np.cumsum([len(st) for st in streams])])
    ctl = np.array(list(woff) + grp + [u for st in streams for u in st] + [nop] * DA, dtype=np.uint32)
    tb.smem_bytes = 8 * smem_doubles + 4 * int(ctl.shape[0])
    if tb.smem_bytes > MAX_SMEM:
        tb.fits = False
        return tb
    lamfrag = lam_pad.reshape(n_chunks, 8, PA_total // 4, 4).transpose(0, 2, 1, 3).reshape(n_chunks, PA_total * 8)

    # stage C: quadrature fragments W[(l0 + lane/4, M), ring = 8c + 4ks + lane%4], the two k-steps of a lane adjacent
    _, Wt = _sf.analysis_tables(s1 + s2, 0, L_out, n_theta, n_phi)       # [(L_out+1)^2, n_theta]
    W_rows = np.zeros((n_tiles, 8, n_rings))
    for ti, (Mi, l0) in enumerate(tiles):
        M = Mi - L_out
        ls = np.arange(l0, min(l0 + 8, L_out + 1))
        W_rows[ti, : ls.shape[0], :n_theta] = Wt[ls * (ls + 1) + M]
    wtfrag = W_rows.reshape(n_tiles, 8, n_chunks, 2, 4).transpose(2, 0, 1, 4, 3).reshape(n_chunks, n_tiles * 64)   # [chunk][tile][lane][k-step]

    qmax = (ell1_max + ell2_max + L_out) // n_phi
    tb.perm1 = (8 * (fields[0]["base"] + fields[0]["perm"])).astype(np.int32)
    tb.perm2 = (8 * (fields[1]["base"] + fields[1]["perm"])).astype(np.int32)
    tb.ctl = ctl.view(np.int32)
    tb.n_ctl = int(ctl.shape[0])
    tb.n_tiles_arg = n_tiles
    tb.lamfrag = np.ascontiguousarray(lamfrag)
    tb.tiles = np.ascontiguousarray(np.array(tiles, dtype=np.int32))
    tb.wtfrag = np.ascontiguousarray(wtfrag)
    tb.cfg = np.array([ell1_max, ell2_max, L_out, n_phi, n_chunks, qmax, szA, offF1, offF2, smem_doubles, nwarps, 0, GM, maxt, 0, 0], dtype=np.int32)
    tb.n_out = (L_out + 1) ** 2
    tb.nwarps = nwarps
    tb.gm = GM
    # algorithmic work per time step
    tb.flops_per_step = float(
        4 * n_theta * (fields[0]["perm"].shape[0] + fields[1]["perm"].shape[0])            # (A) real lambda x complex mode
        + 8 * n_theta * sum(                                                               # (B) complex multiply-add per (m1, m2) pair
            max(0, min(ell1_max, M + ell2_max) - max(-ell1_max, M - ell2_max) + 1) for M in range(-L_out, L_out + 1))
        + 4 * n_theta * tb.n_out                                                           # (C) real weight x complex P_M
    )
    return tb


CLUSTER_GM, CLUSTER_MAXT, CLUSTER_DA = 5, 12, 9   # csrc/product.cu: modes_product_cluster_kernel<5, 12, 9>


def _cluster_tables(s1, ell1_min, ell1_max, s2, ell2_min, ell2_max, n_theta, n_phi, L_out):
    """Tables of the cluster variant (a pair of CTAs per block of time steps; CTA r stages factor r only, convolves and
    integrates its half of the M range): everything is per CTA rank."""
    tb = ProductTables()
    tb.cluster = True
    GM, maxt, DA = CLUSTER_GM, CLUSTER_MAXT, CLUSTER_DA
    n_chunks = (n_theta + 7) // 8
    n_rings = 8 * n_chunks
    n_mout = 2 * L_out + 1
    fields = []
    base = 0
    for s, lmin, lmax in ((s1, ell1_min, ell1_max), (s2, ell2_min, ell2_max)):
        pos, ks, PA, perm, l_lo = _field_layout(lmin, lmax)
        fields.append(dict(s=s, lmin=lmin, lmax=lmax, pos=pos, ks=ks, PA=PA, perm=perm, l_lo=l_lo, base=base))
        base += PA + 4                                      # one all-zero k-step after each factor (no-op steps point there)
    PA_total = base
    n_steps = PA_total // 4
    szA = 8 * (max(f["PA"] for f in fields) + 4)
    nF1 = max(2 * ell1_max + 1, n_mout)
    offF2rel = 64 * nF1 + 64 * (GM + 2)
    bufStride = offF2rel + 64 * (2 * ell2_max + 1) + 64 * (GM - 1)
    smem_doubles = szA + 2 * bufStride
    ng = -(-n_mout // GM)
    gcnt = [(ng + 1) // 2, ng - (ng + 1) // 2]
    g0 = [0, gcnt[0]]
    all_tiles = [(M + L_out, l0) for M in range(-L_out, L_out + 1) for l0 in range(abs(M), L_out + 1, 8)]
    rank_tiles = [[tl for tl in all_tiles if (tl[0] // GM >= g0[r]) and (tl[0] // GM < g0[r] + gcnt[r])] for r in range(2)]
    tiles_r = 8 * maxt
    # control streams per rank (8 synthesis warps)
    lam_pad = np.zeros((n_rings, PA_total))
    ctls = []
    for r, f in enumerate(fields):
        lam = _lambda(f["s"], f["lmax"], n_theta)
        n = f["perm"].shape[0]
        lam_pad[:n_theta, f["base"] + f["perm"]] = lam[:, f["lmin"] ** 2 : f["lmin"] ** 2 + n]
        entry0 = 0 if r == 0 else offF2rel // 64
        tasks = [((f["base"] + int(f["pos"][mi])) // 4, int(f["ks"][mi]), entry0 + mi) for mi in range(2 * f["lmax"] + 1)]
        nop = (f["base"] + f["PA"]) // 4                    # this factor's own zero k-step: inside the CTA's (zeroed) mode tile
        streams = _build_streams(tasks, 8, DA, nop)
        grp = _assign_groups(range(g0[r], g0[r] + gcnt[r]), 8, GM, L_out, ell1_max, ell2_max, n_phi, (ell1_max + ell2_max + L_out) // n_phi)
        woff = 17 + np.concatenate([[0], np.cumsum([len(st) for st in streams])])
        ctls.append(list(woff) + grp + [u for st in streams for u in st] + [nop] * DA)
    n_ctl_r = max(len(c) for c in ctls)
    ctl = np.array([c + [c[-1]] * (n_ctl_r - len(c)) for c in ctls], dtype=np.uint32)
    tb.fits = (
        max(gcnt) <= 8 and max(len(t) for t in rank_tiles) <= tiles_r and 8 * smem_doubles + 4 * n_ctl_r <= MAX_SMEM
        and n_steps < 65535 and offF2rel // 64 + 2 * ell2_max + 1 < 32768 and n_theta >= 2
    )
    tb.smem_bytes = 8 * smem_doubles + 4 * n_ctl_r
    if not tb.fits:
        return tb
    lamfrag = lam_pad.reshape(n_chunks, 8, PA_total // 4, 4).transpose(0, 2, 1, 3).reshape(n_chunks, PA_total * 8)
    _, Wt = _sf.analysis_tables(s1 + s2, 0, L_out, n_theta, n_phi)
    W_rows = np.zeros((2, tiles_r, 8, n_rings))
    tiles = np.zeros((2, tiles_r, 2), dtype=np.int32)
    tiles[:, :, 1] = L_out + 1                                # empty tiles: no row is stored
    for r in range(2):
        for ti, (Mi, l0) in enumerate(rank_tiles[r]):
            ls = np.arange(l0, min(l0 + 8, L_out + 1))
            W_rows[r, ti, : ls.shape[0], :n_theta] = Wt[ls * (ls + 1) + Mi - L_out]
            tiles[r, ti] = (Mi, l0)
    wtfrag = W_rows.reshape(2, tiles_r, 8, n_chunks, 2, 4).transpose(3, 0, 1, 2, 5, 4).reshape(n_chunks, 2 * tiles_r * 64)
    qmax = (ell1_max + ell2_max + L_out) // n_phi
    tb.perm1 = (8 * fields[0]["perm"]).astype(np.int32)
    tb.perm2 = (8 * fields[1]["perm"]).astype(np.int32)
    tb.ctl = np.ascontiguousarray(ctl).view(np.int32).reshape(-1)
    tb.n_ctl = int(n_ctl_r)
    tb.lamfrag = np.ascontiguousarray(lamfrag)
    tb.tiles = np.ascontiguousarray(tiles.reshape(-1, 2))
    tb.n_tiles_arg = tiles_r
    tb.wtfrag = np.ascontiguousarray(wtfrag)
    tb.cfg = np.array([ell1_max, ell2_max, L_out, n_phi, n_chunks, qmax, szA, bufStride, offF2rel, smem_doubles, 16, 0, GM, maxt, 0,
                       1, fields[1]["base"] // 4, g0[0], gcnt[0], g0[1], gcnt[1]], dtype=np.int32)
    tb.n_out = (L_out + 1) ** 2
    tb.nwarps, tb.gm = 16, GM
    tb.flops_per_step = product_tables(s1, ell1_min, ell1_max, s2, ell2_min, ell2_max, n_theta, n_phi, L_out, 0).flops_per_step \
        if product_tables(s1, ell1_min, ell1_max, s2, ell2_min, ell2_max, n_theta, n_phi, L_out, 0).fits else 0.0
    return tb


THETA_FSTRIDE, THETA_DA, THETA_MAX_SMEM = 66, 9, 113 * 1024   # csrc/product.cu: theta_synth_kernel<9>


@lru_cache(maxsize=16)
def theta_tables(s, ell_min, ell_max, n_theta):
    """Tables of scrib200_theta_synth (the theta stage of the separable salm2map): the single-field subset of the product
    tables - m-major mode permutation, lambda fragments per ring chunk, the eight warps' control streams."""
    tb = ProductTables()
    DA = THETA_DA
    n_chunks = (n_theta + 7) // 8
    n_rings = 8 * n_chunks
    pos, ks, PA, perm, l_lo = _field_layout(ell_min, ell_max)
    nm = 2 * ell_max + 1
    n_steps = PA // 4 + 1                                   # + one all-zero k-step: the target of no-op steps
    szA = 8 * (PA + 4)
    smem_doubles = szA + nm * THETA_FSTRIDE
    tasks = [(int(pos[mi]) // 4, int(ks[mi]), mi) for mi in range(nm)]
    streams = _build_streams(tasks, 8, DA, n_steps - 1)
    woff = 9 + np.concatenate([[0], np.cumsum([len(st) for st in streams])])
    ctl = np.array(list(woff) + [u for st in streams for u in st] + [n_steps - 1] * DA, dtype=np.uint32)
    tb.smem_bytes = 8 * smem_doubles + 4 * int(ctl.shape[0])
    tb.fits = tb.smem_bytes <= THETA_MAX_SMEM and n_steps < 65535 and n_theta >= 2
    if not tb.fits:
        return tb
    lam = _lambda(s, ell_max, n_theta)
    lam_pad = np.zeros((n_rings, PA + 4))
    lam_pad[:n_theta, perm] = lam[:, ell_min**2 :]
    tb.lamfrag = np.ascontiguousarray(lam_pad.reshape(n_chunks, 8, PA // 4 + 1, 4).transpose(0, 2, 1, 3).reshape(n_chunks, PA * 8 + 32))
    tb.perm = (8 * perm).astype(np.int32)
    tb.ctl = ctl.view(np.int32)
    tb.n_ctl = int(ctl.shape[0])
    tb.cfg = np.array([n_theta, nm, n_chunks, szA, smem_doubles], dtype=np.int32)
    tb.nm = nm
    return tb


QUAD_MAXT = 21   # csrc/product.cu: theta_quad_kernel<21>


@lru_cache(maxsize=16)
def quad_tables(s, ell_min, ell_max, n_theta, n_phi):
    """Tables of scrib200_theta_quad (the theta stage of the separable map2salm): output tiles of 8 consecutive l per M
    from max(|M|, ell_min), and the quadrature weights 2 pi q_j sY_lM(theta_j, 0) as DMMA fragments per ring chunk."""
    tb = ProductTables()
    L = ell_max
    n_chunks = (n_theta + 7) // 8
    n_rings = 8 * n_chunks
    tiles = [(M + L, l0) for M in range(-L, L + 1) for l0 in range(max(abs(M), ell_min), L + 1, 8)]
    n_tiles = 8 * QUAD_MAXT
    tb.fits = len(tiles) <= n_tiles and n_theta >= 2 and (2 * L + 1) * THETA_FSTRIDE * 8 <= 100 * 1024
    if not tb.fits:
        return tb
    _, Wt = _sf.analysis_tables(s, 0, L, n_theta, n_phi)
    W_rows = np.zeros((n_tiles, 8, n_rings))
    for ti, (Mi, l0) in enumerate(tiles):
        ls = np.arange(l0, min(l0 + 8, L + 1))
        W_rows[ti, : ls.shape[0], :n_theta] = Wt[ls * (ls + 1) + Mi - L]
    tiles += [(0, L + 1)] * (n_tiles - len(tiles))
    tb.wtfrag = np.ascontiguousarray(W_rows.reshape(n_tiles, 8, n_chunks, 2, 4).transpose(2, 0, 1, 4, 3).reshape(n_chunks, n_tiles * 64))
    tb.tiles = np.ascontiguousarray(np.array(tiles, dtype=np.int32))
    tb.n_out = (L + 1) ** 2 - ell_min**2
    tb.cfg = np.array([n_theta, 2 * L + 1, n_chunks, tb.n_out, ell_min**2, L], dtype=np.int32)
    return tb
