"""Batch sharding across the GPUs of one box.

Frame pairs are independent units: rank r of G owns a contiguous slice of the batch, runs the
alignment kernel on it with no data-path collective, and the per-pair results (7-double poses,
< 0.5 KB per pair with H) are gathered with one all_gather.  torch.distributed is the plumbing
(NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice of rank `rank`; the first n_items % world ranks get one extra item."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_rows(local: np.ndarray, n_items: int, device: torch.device | str = "cpu") -> np.ndarray:
    """all_gather of per-item result rows (e.g. [n_local, 7] poses) into the global [n_items, ...] array,
    in batch order, on every rank.  Uneven shards are padded to the largest shard for the collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    cap = max(e - b for b, e in sizes)
    row = local.shape[1:]
    buf = torch.zeros((cap,) + row, dtype=torch.from_numpy(local).dtype, device=device)
    buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    out = torch.empty((world * cap,) + row, dtype=buf.dtype, device=device)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape((world, cap) + row)
    return np.concatenate([out[r, : e - b] for r, (b, e) in enumerate(sizes)], axis=0)
