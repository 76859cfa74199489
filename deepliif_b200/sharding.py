"""Tile sharding across one-process-per-GPU ranks.

Tiles are independent given per-sample normalisation statistics, so inference shards with no data-path collective:
rank r takes tiles r, r+W, r+2W, ... and holds a full replica of the generators.  The only communication is the
gather of uint8 results for stitching on rank 0 (this *replaces* the reference's net-group placement,
deepliif/models/__init__.py:201-211, which serialises tiles and cannot use more than 5 GPUs)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """Indices owned by `rank` (round-robin: consecutive tiles of a row land on different GPUs)."""
    return list(range(rank, n_items, world))


def shard(items, rank, world):
    return items[rank::world]


def gather_to_rank0(local, n_items, group=None, to_numpy=True):
    """local: uint8 tensor/array [n_local, ...] holding this rank's results in shard order.  Returns on rank 0 the
    full [n_items, ...] array in the original tile order (None elsewhere).  Works on gloo (CPU) and nccl (CUDA).
    to_numpy=False keeps the result a tensor on the gather's device (rank 0 of an NCCL group: its GPU)."""
    if not (dist.is_available() and dist.is_initialized()):
        if not to_numpy:
            return local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
        return np.asarray(local.cpu() if isinstance(local, torch.Tensor) else local)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    backend = dist.get_backend(group)
    if backend == "nccl" and not t.is_cuda:
        t = t.cuda()
    per_rank_max = (n_items + world - 1) // world
    pad = torch.zeros((per_rank_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]].copy_(t)
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    if not to_numpy:
        out_t = torch.empty((n_items,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        for r in range(world):
            idx = shard_indices(n_items, r, world)
            if idx:
                out_t[torch.tensor(idx, device=t.device)] = bufs[r][: len(idx)]
        return out_t
    out = np.empty((n_items,) + tuple(t.shape[1:]), dtype=np.asarray(pad.cpu()).dtype)
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        out[idx] = bufs[r][: len(idx)].cpu().numpy()
    return out


def max_over_ranks(value, device="cpu", group=None):
    """Timing helper: the slowest rank defines the step time."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
