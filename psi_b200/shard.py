"""Read sharding across the GPUs of one box and the (only) collective of the path.

Seed finding shards by read: every rank gets a contiguous range of reads holding about the
same number of bases, keeps the whole graph + index, and produces records whose read ids are disjoint
from every other rank's, so the concatenation of the per-rank outputs in rank order IS the global
result.  Nothing is exchanged on the data path; what is reduced afterwards is bookkeeping: counts and a
hits-per-read histogram (`torch.distributed` all-reduce: NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

HIST_BINS = 64


def shard_bounds(read_ptr: np.ndarray, world: int) -> np.ndarray:
    """Read indices b[0..world]: rank i owns reads [b[i], b[i+1]); ranges hold ~equal base counts."""
    read_ptr = np.asarray(read_ptr, dtype=np.uint64)
    n = len(read_ptr) - 1
    if world < 1:
        raise ValueError("world must be >= 1")
    total = int(read_ptr[-1] - read_ptr[0])
    targets = (np.arange(1, world, dtype=np.float64) * (total / world) + float(read_ptr[0])).astype(np.uint64)
    cuts = np.searchsorted(read_ptr, targets, side="left") if world > 1 else np.zeros(0, np.int64)
    b = np.concatenate([[0], np.minimum(cuts, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(b)


def shard_of(read_ptr: np.ndarray, bases: np.ndarray, rank: int, world: int, first_read_id: int = 0):
    """(read_ptr, bases, first_read_id) of this rank's shard; read ids stay global (sequence.hpp:1616)."""
    b = shard_bounds(read_ptr, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    sub_ptr = np.asarray(read_ptr[lo:hi + 1], dtype=np.uint64) - np.uint64(read_ptr[lo])
    sub_bases = np.asarray(bases)[int(read_ptr[lo]):int(read_ptr[hi])]
    return sub_ptr, sub_bases, first_read_id + lo


def hits_per_read_histogram(read_ids: np.ndarray, n_reads: int, first_read_id: int, bins: int = HIST_BINS) -> np.ndarray:
    """hist[h] = number of reads of this shard with h hits (last bin: >= bins-1)."""
    per_read = np.bincount(np.asarray(read_ids, dtype=np.int64) - first_read_id, minlength=n_reads)[:n_reads]
    return np.bincount(np.minimum(per_read, bins - 1), minlength=bins).astype(np.int64)


def all_reduce_counts(counts: dict, hist: np.ndarray, device=None):
    """Sum `counts` (name -> int) and `hist` over all ranks.  No-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    names = sorted(counts)
    t = torch.tensor([int(counts[k]) for k in names] + [int(x) for x in hist], dtype=torch.int64, device=device or "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    v = t.cpu().tolist()
    return dict(zip(names, v[:len(names)])), np.array(v[len(names):], dtype=np.int64)
