"""Multi-GPU layout of the path: independent environments (bodies) are partitioned across ranks; there is NO
data-path collective. The only exchange is the validation gather C1: every rank's sfh_info records
and a 64-bit order-independent checksum, outside the timed region (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition [lo, hi) of n_units over `world` ranks; sizes differ by at most one."""
    base, rem = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def checksum(infos: np.ndarray) -> int:
    """Order-independent 64-bit checksum of a set of sfh_info records: wrapping sum of a per-record hash."""
    if infos.shape[0] == 0:
        return 0
    words = np.ascontiguousarray(infos).view(np.uint64).reshape(infos.shape[0], -1)
    with np.errstate(over="ignore"):
        h = np.full(infos.shape[0], 0xCBF29CE484222325, dtype=np.uint64)
        for c in range(words.shape[1]):
            h = (h ^ words[:, c]) * np.uint64(0x100000001B3)
            h ^= h >> np.uint64(29)
        return int(h.sum(dtype=np.uint64))


def gather_infos(local: np.ndarray, group=None, device=None) -> tuple[np.ndarray, list[int]]:
    """All-gather per-rank sfh_info arrays (possibly of different lengths) and their checksums.
    Returns (all records in rank order, per-rank checksums). Uses the default process group's backend:
    tensors live on `device` (cuda for NCCL over NVLink, cpu for gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    n_local = torch.tensor([local.shape[0], checksum(local) & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=dev)
    meta = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(meta, n_local, group=group)
    counts = [int(m[0]) for m in meta]
    sums = [int(m[1]) for m in meta]
    rec = local.dtype.itemsize
    cap = max(counts) if counts else 0
    buf = torch.zeros(max(cap * rec, 1), dtype=torch.uint8, device=dev)
    if local.shape[0]:
        buf[: local.shape[0] * rec] = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).reshape(-1).copy()).to(dev)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = [p[: c * rec].cpu().numpy().view(local.dtype) for p, c in zip(parts, counts)]
    return (np.concatenate(out) if out else local[:0]), sums
