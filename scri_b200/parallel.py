"""Multi-GPU sharding of the transformation path (one process per GPU, torch.distributed).

The path shards two ways:

* by waveform index - fully independent units, no data-path collective; `shard_range` gives each rank a
  contiguous slice of the batch and `gather_batch` is the optional final all_gather;
* by time - pointwise stages need nothing; the spline stages couple neighbouring samples only through the
  tridiagonal inverse, which decays like 0.268^k, so each rank takes its block plus a HALO-sample halo of
  *input modes* (`time_shard_with_halo`), transforms locally and keeps the outputs that fall in its own
  block (`owned_output_mask`).  The halo rows come from the neighbours by point-to-point exchange
  (`exchange_halos`) - 2 * HALO * n_modes * 16 bytes per boundary - and the results are all_gathered.

The host logic here is backend-agnostic: NCCL on GPUs, gloo in the CPU tests.
"""
import numpy as np

HALO = 64


def shard_range(n_items, rank, world_size):
    """Contiguous [lo, hi) slice of `n_items` owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def time_shard_with_halo(n_times, rank, world_size, halo=HALO):
    """(lo, hi, lo_h, hi_h): owned block [lo, hi) and the block with halos [lo_h, hi_h)."""
    lo, hi = shard_range(n_times, rank, world_size)
    return lo, hi, max(0, lo - halo), min(n_times, hi + halo)


def owned_output_mask(uprm, t_global_first_owned, t_global_first_next, gamma, time_translation, is_last):
    """Outputs u'_i = (t_i - dt)/gamma belong to the rank that owns input sample i."""
    u_lo = (1 / gamma) * (t_global_first_owned - time_translation)
    if is_last:
        return uprm >= u_lo
    u_hi = (1 / gamma) * (t_global_first_next - time_translation)
    return (uprm >= u_lo) & (uprm < u_hi)


def exchange_halos(local, halo, group=None):
    """Send the first/last `halo` rows of `local` ([n_local, ...] tensor) to the previous/next rank.

    Returns (from_prev, from_next); None at the ends of the chain.  Works with NCCL (CUDA tensors) and gloo.
    """
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    h = min(halo, local.shape[0])
    from_prev = torch.empty((h,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) if rank > 0 else None
    from_next = torch.empty((h,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) if rank < world - 1 else None
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, local[:h].contiguous(), rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, from_prev, rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, local[-h:].contiguous(), rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, from_next, rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return from_prev, from_next


def gather_batch(local, counts, group=None):
    """all_gather of per-rank row blocks of unequal length `counts` (rows along dim 0)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    nmax = max(counts)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if pad.is_complex():
        buf = torch.view_as_real(pad).contiguous()
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf, group=group)
        outs = [torch.view_as_complex(o) for o in outs]
    else:
        outs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(outs, counts)], dim=0)


def transform_batch(plan, t, data_batch):
    """Run one TransformPlan over a local shard `data_batch[B_local, N, n]` (device tensors): (u', modes'[B_local, N', n'])."""
    return plan.run_batch(t, data_batch)
