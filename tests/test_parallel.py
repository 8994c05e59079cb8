"""world_size-2 gloo tests of the sharding logic (CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scri_b200 import parallel


def test_shard_ranges_cover_everything():
    for n, w in [(10, 3), (4096, 8), (7, 8), (100001, 4)]:
        spans = [parallel.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    lo, hi, lo_h, hi_h = parallel.time_shard_with_halo(1000, 1, 4, halo=64)
    assert (lo, hi, lo_h, hi_h) == (250, 500, 186, 564)
    assert parallel.time_shard_with_halo(1000, 0, 4)[2] == 0 and parallel.time_shard_with_halo(1000, 3, 4)[3] == 1000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_times, n = 40, 5
        full = torch.arange(n_times * n, dtype=torch.float64).reshape(n_times, n)
        full = torch.complex(full, -full)
        lo, hi = parallel.shard_range(n_times, rank, world)
        local = full[lo:hi].clone()
        prev, nxt = parallel.exchange_halos(local, 4)
        ok = True
        if rank > 0:
            ok &= bool(torch.equal(prev, full[lo - 4 : lo]))
        else:
            ok &= prev is None
        if rank < world - 1:
            ok &= bool(torch.equal(nxt, full[hi : hi + 4]))
        else:
            ok &= nxt is None
        # unequal shards gathered back in order
        counts = [parallel.shard_range(n_times + 1, r, world)[1] - parallel.shard_range(n_times + 1, r, world)[0] for r in range(world)]
        full2 = torch.arange((n_times + 1) * n, dtype=torch.float64).reshape(n_times + 1, n)
        lo2, hi2 = parallel.shard_range(n_times + 1, rank, world)
        gathered = parallel.gather_batch(full2[lo2:hi2].clone(), counts)
        ok &= bool(torch.equal(gathered, full2))
        gathered_c = parallel.gather_batch(torch.complex(full2, full2)[lo2:hi2].clone(), counts)
        ok &= bool(torch.equal(gathered_c, torch.complex(full2, full2)))
        # time-sharded ownership of outputs: every output belongs to exactly one rank
        t = np.linspace(0.0, 10.0, n_times)
        gamma, tt = 1.0007, 0.3
        uprm = (1 / gamma) * (t - tt)
        uprm = uprm[3:-2]
        lo, hi = parallel.shard_range(n_times, rank, world)
        mask = parallel.owned_output_mask(uprm, t[lo], t[hi] if hi < n_times else None, gamma, tt, rank == world - 1)
        cnt = torch.tensor([int(mask.sum())])
        dist.all_reduce(cnt)
        ok &= int(cnt) == uprm.shape[0]
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_halo_exchange_and_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_transform_halo_grows_with_boost_and_time():
    """Halo of the time-sharded transform: spline decay + the distance between an output time and the input samples the
    grid points need for it - constant for supertranslations, beta |t| / dt under a boost."""
    from types import SimpleNamespace

    from scri_b200 import parallel

    G = 50
    rng = np.random.default_rng(3)
    alpha = rng.uniform(-0.3, 0.3, G)
    still = SimpleNamespace(kconformal=np.ones(G), alpha=alpha, gamma=1.0, time_translation=0.1)
    h0 = parallel.transform_halo(still, 0.0, 1000.0, 0.1)
    assert h0 == int(np.ceil(np.abs(alpha - 0.1).max() / 0.1)) + parallel.SPLINE_DECAY_ROWS + 2
    assert parallel.transform_halo(still, 5000.0, 9000.0, 0.1) == h0                  # no drift without a boost
    beta = 0.03
    vr = rng.uniform(-beta, beta, G)
    gamma = 1 / np.sqrt(1 - beta**2)
    boosted = SimpleNamespace(kconformal=1 / (gamma * (1 - vr)), alpha=alpha, gamma=gamma, time_translation=0.1)
    h1 = parallel.transform_halo(boosted, 0.0, 1000.0, 0.1)
    h2 = parallel.transform_halo(boosted, 0.0, 2000.0, 0.1)
    assert h1 > h0 + 0.9 * np.abs(vr).max() * 1000.0 / 0.1 - 10 and h2 > 1.9 * (h1 - h0)
