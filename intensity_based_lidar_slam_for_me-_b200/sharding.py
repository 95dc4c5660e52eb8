"""Host-side logic of the one sharded path: a ScanContext keyframe database split into contiguous id
ranges, one per rank; per query every rank scores its shard and keeps a local top-k, the ranks exchange
k x (f64 distance, i32 id, i32 shift) with ONE all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests) and merge
identically.  The message is k*16 B per rank: latency-bound, nothing to overlap."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) of rank `rank` when n keyframes are split over `world` ranks (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_topk(dist, ids, shifts, device=None):
    """all_gather of one rank's top-k triplet through torch.distributed (must be initialised).  Returns the
    concatenated (dist, ids, shifts) of all ranks as numpy arrays, identical on every rank."""
    import torch
    import torch.distributed as td
    k = len(dist)
    dev = device if device is not None else "cpu"
    # one message: k doubles followed by 2k int32 viewed as k doubles
    pack = torch.empty(2 * k, dtype=torch.float64, device=dev)
    pack[:k] = torch.as_tensor(np.asarray(dist, np.float64), device=dev)
    ii = np.concatenate([np.asarray(ids, np.int32), np.asarray(shifts, np.int32)])
    pack[k:] = torch.as_tensor(ii.view(np.float64), device=dev)
    out = [torch.empty_like(pack) for _ in range(td.get_world_size())]
    td.all_gather(out, pack)
    got = torch.stack(out).cpu().numpy()
    d = got[:, :k].reshape(-1)
    rest = np.ascontiguousarray(got[:, k:]).view(np.int32).reshape(len(out), 2 * k)
    return d, rest[:, :k].reshape(-1).copy(), rest[:, k:].reshape(-1).copy()
