"""Multi-GPU plumbing: meshes are independent units, so the forward shards them across ranks
with no data-path collective; training adds exactly one exchange step, a
single flat-buffer all-reduce of the parameter gradients (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  The reference has no distributed code; this is a new component."""
from __future__ import annotations

import torch
import torch.distributed as dist


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPU cores NVML reports as local to GPU ``local_rank`` (its NUMA node).

    Call BEFORE allocating pinned host buffers: pinned pages are placed on the node of the allocating thread, and
    on an 8-GPU box (GPUs 0-3 on socket 0, 4-7 on socket 1) a buffer on the wrong socket makes every H2D/D2H copy
    cross the inter-socket link -- round 1's end-to-end scaling fell to 0.56 at 8 GPUs for exactly that reason.
    Returns the list of CPUs bound to, or None when the topology is unavailable (never raises)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                idx = int(ids[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def mesh_cost(V, K, C, nnz_per_row=7, bytes_per_el=4):
    """Algorithmic HBM bytes of one block forward on one mesh."""
    return V * (bytes_per_el * (5 * C + 2 * K) + 12 * nnz_per_row + 8)


def shard_meshes(costs, world_size):
    """Greedy longest-processing-time assignment of mesh indices to ranks.

    Returns a list (len world_size) of index lists; deterministic, every index exactly once."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda q: (loads[q], q))
        shards[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in shards]


def allreduce_gradients(params, n_global_meshes=None, group=None):
    """One all-reduce(SUM) over a single flat fp32 buffer holding every parameter gradient
    (<= 14 MB for the largest reference model: latency-bound, so no bucketing).  Divides by
    ``n_global_meshes`` when given (gradient of the mean loss over the global batch)."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    dev = params[0].device
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        if p.grad is not None:
            flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
        off += p.numel()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if n_global_meshes:
        flat.div_(float(n_global_meshes))
    off = 0
    for p in params:
        g = flat[off:off + p.numel()].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += p.numel()
