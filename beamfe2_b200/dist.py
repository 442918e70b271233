"""Multi-GPU: the batch of independent structure variants is cut into contiguous per-rank slices; the
topology plan is replicated; there is NO data-path collective.  The only exchange is one final
gather of small per-structure summaries (torch.distributed: NCCL on GPUs, gloo in CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of rank `rank`; sizes differ by at most one, union is [0, total)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; no-op for a single process.
    Returns (rank, local_rank, world)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29511')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def bind_to_gpu_numa_node(device_index: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE pinned host buffers are
    allocated (first touch places them on that node).  With one process per GPU the host side of the
    end-to-end path (H2D / D2H through pinned memory) otherwise crosses the socket interconnect.
    Returns the node id or None when the topology cannot be read (then nothing is changed)."""
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
    except Exception:
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
        except Exception:
            return None
    try:
        bus = str(bus).lower()
        if bus.count(':') == 2 and len(bus.split(':')[0]) == 8:
            bus = bus[4:]                                   # 00000000:1b:00.0 -> 0000:1b:00.0
        with open('/sys/bus/pci/devices/%s/numa_node' % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        return None
    return None


def summarize(u, lam, info, n_freq=3):
    """Per-structure summary [B, 2 + n_freq]: info, max |u| over load cases and DOFs, first circular
    frequencies.  ~40 B per structure -- this is all that ever crosses NVLink."""
    B = info.shape[0]
    cols = [info.to(torch.float64).reshape(B, 1)]
    if u is not None:
        cols.append(u.abs().amax(dim=(1, 2)).reshape(B, 1))
    else:
        cols.append(torch.zeros((B, 1), dtype=torch.float64, device=info.device))
    if lam is not None:
        cols.append(torch.sqrt(torch.clamp(lam[:, :n_freq], min=0.0)))
    return torch.cat(cols, dim=1).contiguous()


def gather_summaries(local: torch.Tensor, total: int):
    """All ranks call; every rank gets the [total, k] summary table in global batch order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(total, r, world) for r in range(world)]
    k = local.shape[1]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn, k), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
