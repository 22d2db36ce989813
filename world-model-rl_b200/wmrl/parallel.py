"""Data parallelism over start states: one process per GPU,
weights replicated, the batch split contiguously, no exchange on the rollout
itself.  The only collectives are one flat all-reduce of the actor (or RSSM)
gradients after the backward and a scalar metric reduction."""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n rows owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(t: torch.Tensor, rank: int, world: int, dim: int = 0) -> torch.Tensor:
    lo, hi = shard_range(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = True,
                        weight: Optional[float] = None) -> int:
    """One all-reduce over a flat fp32 buffer of every existing .grad.
    `weight` rescales the local gradient first (rows_local / rows_global for a
    mean loss over unevenly sharded rows).  Returns the number of elements sent."""
    grads: List[torch.Tensor] = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized():
        return 0
    world = dist.get_world_size(group)
    flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
    if weight is not None:
        flat.mul_(weight)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average and weight is None:
        flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel()


def reduce_scalar(value: float, device, op=dist.ReduceOp.SUM, group=None) -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op, group=group)
    return float(t.item())


def bind_to_gpu_numa(device_index: int) -> bool:
    """Pin the calling process to the CPUs NVML reports as local to GPU `device_index` (one process per GPU).
    Host buffers allocated afterwards (pinned staging memory) are first-touched on that NUMA node, so the per-step
    host->device copies of 8 ranks do not all cross the socket interconnect.  Best effort: returns False when NVML
    or the affinity call is unavailable."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False
