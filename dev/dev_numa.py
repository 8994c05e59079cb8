"""Developer script (torchrun, one rank per GPU): where do the host buffers of each rank live relative to its GPU, and what
does concurrent H2D + D2H traffic of all ranks achieve (a) as placed by default, (b) with the rank bound to the CPUs nvml
reports as local to its GPU before the buffers are allocated."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
if rank == 0:
    for cmd in ("nvidia-smi topo -m", "lscpu | grep -i 'numa\\|socket\\|model name'", "cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective",
                "cat /sys/devices/system/node/online", "grep -i 'MemTotal\\|MemFree' /sys/devices/system/node/node*/meminfo"):
        print("$", cmd); print(subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout)
    sys.stdout.flush()
allowed = sorted(os.sched_getaffinity(0))
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(local)
words = pynvml.nvmlDeviceGetCpuAffinity(h, 16)
gpu_cpus = [64 * w + b for w, x in enumerate(words) for b in range(64) if (x >> b) & 1]
try:
    memw = pynvml.nvmlDeviceGetMemoryAffinity(h, 4, pynvml.NVML_AFFINITY_SCOPE_NODE)
    nodes = [64 * w + b for w, x in enumerate(memw) for b in range(64) if (x >> b) & 1]
except Exception as e:
    nodes = repr(e)
for r in range(world):
    barrier()
    if r == rank:
        print(f"rank {rank}: allowed cpus {allowed[0]}..{allowed[-1]} ({len(allowed)}), on cpu {os.sched_getcpu() if hasattr(os,'sched_getcpu') else '?'}; gpu-local cpus {gpu_cpus[0] if gpu_cpus else None}..{gpu_cpus[-1] if gpu_cpus else None} ({len(gpu_cpus)}), gpu numa node(s) {nodes}; local&allowed {len(set(gpu_cpus)&set(allowed))}")
        sys.stdout.flush()

def node_of(tensor):
    """NUMA node of the first page of a host tensor (move_pages with a null target = query)."""
    import ctypes
    libc = ctypes.CDLL(None, use_errno=True)
    page = ctypes.c_void_p(tensor.data_ptr() & ~4095)
    status = ctypes.c_int(-1)
    rc = libc.syscall(279, 0, 1, ctypes.byref(page), None, ctypes.byref(status), 0)   # __NR_move_pages on x86_64
    return status.value if rc == 0 else f"errno {ctypes.get_errno()}"

def measure(tag):
    n = 124_000_000
    src = torch.empty(n, dtype=torch.uint8, pin_memory=True); src.fill_(1)
    dst = torch.empty(n, dtype=torch.uint8, pin_memory=True); dst.fill_(2)
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.ones(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for rep in range(2):
            barrier()
            t0 = time.perf_counter()
            for _ in range(10):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1): d_in.copy_(src, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2): dst.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 10
        res[mode] = dt * 1e3
    for r in range(world):
        barrier()
        if r == rank:
            print(f"[{tag}] rank {rank}: buffers on node {node_of(src)}/{node_of(dst)}; 124 MB h2d {res['h2d']:.2f} ms ({0.124/res['h2d']*1e3:.1f} GB/s), d2h {res['d2h']:.2f} ms, both {res['both']:.2f} ms")
            sys.stdout.flush()
measure("default")
both = sorted(set(gpu_cpus) & set(allowed))
if both:
    os.sched_setaffinity(0, both)
measure("bound to gpu-local cpus" if both else "no gpu-local cpu allowed")
if world > 1:
    dist.destroy_process_group()
