"""Frame-pair sharding across GPUs (one process per GPU, no collective on the data path).

The reference has no multi-GPU inference; BASELINE.json's north_star shards independent frame pairs
across the GPUs of one box.  Each rank gets a contiguous range of pair indices; results are concatenated in index
order.  torch.distributed is only used for the final (optional) ordered gather (one all_gather of the result tensors) and
for timing barriers."""
from __future__ import annotations

import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index):
    """Pin the calling process to the CPU cores closest to GPU `device_index` (NVML's ideal affinity mask).  Call it
    before allocating pinned frame buffers: with one process per GPU the host side of every PCIe copy then stays on
    the GPU's own NUMA node.  Returns the number of cores bound, or 0 when NVML / affinity calls are unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def shard_range(n_items, rank, world_size):
    """Contiguous [lo, hi) range of `rank`: sizes differ by at most one, earlier ranks take the remainder."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def run_sharded(infer, frames, timesteps=(0.5,), gather=True):
    """Interpolate every adjacent pair of `frames` (a list/tensor of (C,H,W) or (1,C,H,W) frames) at each timestep.

    `infer(img0, img1, t)` is Model.inference-like.  Pairs are sharded contiguously over the ranks of the default
    process group (or run locally when torch.distributed is not initialised).  Returns, on every rank when
    `gather` is set, the list of outputs ordered by (pair index, timestep index); otherwise the local slice."""
    n_pairs = len(frames) - 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(n_pairs, rank, world)
    local = []
    for i in range(lo, hi):
        a, b = frames[i], frames[i + 1]
        if a.dim() == 3:
            a, b = a.unsqueeze(0), b.unsqueeze(0)
        for t in timesteps:
            local.append(infer(a, b, t))
    if not gather or world == 1:
        return local
    # ordered gather as ONE tensor collective on the outputs' own device (NCCL over NVLink for CUDA tensors, gloo for
    # CPU tensors): every rank contributes a (max_count, ...) block, padded by at most one item, then trims by the
    # ranks' known counts.  No pickling, no host round trip for device tensors.
    per_pair = len(timesteps)
    counts = [(shard_range(n_pairs, r, world)[1] - shard_range(n_pairs, r, world)[0]) * per_pair for r in range(world)]
    cmax = max(counts)
    if local:
        item = local[0]
        shape, dtype, device = tuple(item.shape), item.dtype, item.device
    else:   # a rank without pairs still takes part in the collective: learn the item shape from rank 0
        shape = dtype = device = None
    meta = [None]
    if rank == 0:
        meta = [(shape, dtype)]
    dist.broadcast_object_list(meta, src=0)
    shape, dtype = meta[0]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    if dist.get_backend() != "nccl":
        device = torch.device("cpu")   # gloo: the collective runs on host tensors
    block = torch.zeros((cmax,) + shape, dtype=dtype, device=device)
    for k, x in enumerate(local):
        block[k].copy_(x)
    blocks = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(blocks, block)
    out = []
    for r in range(world):
        out.extend(blocks[r][k] for k in range(counts[r]))
    return out
