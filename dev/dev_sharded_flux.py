"""Developer script (torchrun, NCCL): BASELINE config 5 - fluxes of a 1e6-step l<=16 strain series sharded by time over
the ranks: one NCCL halo exchange for the spline derivative, everything else pointwise.  Prints time and throughput."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
world, rank, lrank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lrank)
dist.init_process_group("nccl" if world > 1 else "gloo", device_id=torch.device("cuda", lrank) if world > 1 else None,
                        init_method=None if "MASTER_ADDR" in os.environ else "tcp://127.0.0.1:29531", rank=rank, world_size=world)
from scri_b200 import parallel, _lib
N, LMIN, LMAX = 1_000_000, 2, 16
n = LMAX * (LMAX + 2) - LMIN**2 + 1
lo, hi = parallel.shard_range(N, rank, world)
g = torch.Generator(device="cuda").manual_seed(0)            # same series on every rank, each keeps its block
w = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 0.45 + 0.05
c = torch.randn(n, dtype=torch.complex128, device="cuda", generator=g)
t = torch.linspace(0.0, 1e5, N, dtype=torch.float64, device="cuda")[lo:hi].clone()
data = c[None, :] * torch.exp(1j * w[None, :] * t[:, None])
for it in range(4):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); E, p, J = parallel.sharded_fluxes(t, data, LMIN, LMAX); e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
chk = torch.stack([E.sum(), p.abs().sum(), J.abs().sum()]).double()
dist.all_reduce(chk)
if rank == 0:
    print(f"world {world}: sharded_fluxes over N = {N}, n = {n}: {float(ms):.3f} ms (max over ranks), {n * N / float(ms) / 1e6:.2f} G mode-timesteps/s, checksums {[f'{float(x):.12e}' for x in chk]}")
dist.destroy_process_group()
