"""Multi-GPU plumbing of the PSRS job: replicas shard across ranks, the bucketed log is replicated, and the
only collective on the simulation path is one small sum of the learning-curve accumulators at the end.
Ingest of a host-resident log may use a second one: every rank uploads 1/world of the rows over its own PCIe link and
the shards are all-gathered over NVLink, so the host side moves the log once instead of ``world`` times.

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is plumbing; nothing on the step path communicates.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def replica_shard(n_total: int, rank: int, world: int):
    """Contiguous block of global replica ids owned by ``rank``: returns ``(replica_id0, n_local)``.
    Blocks differ by at most one replica; ids key the Philox streams and ring offsets, so a job gives the same
    per-replica results for any ``world``."""
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    base, extra = divmod(int(n_total), int(world))
    n_local = base + (1 if rank < extra else 0)
    id0 = rank * base + min(rank, extra)
    return id0, n_local


def row_shard(n_rows: int, rank: int, world: int):
    """Rows ``[lo, hi)`` of a log that ``rank`` uploads in a sharded ingest: equal shards (``all_gather_into_tensor`` needs
    them), so ``n_rows`` must be a multiple of ``world``."""
    if n_rows % world:
        raise ValueError("sharded ingest needs n_rows to be a multiple of the world size")
    per = n_rows // world
    return rank * per, (rank + 1) * per


def all_gather_columns(shards: dict, full: dict) -> None:
    """``full[k]`` (all rows, on this rank's device) from every rank's ``shards[k]`` (its ``row_shard`` of column ``k``):
    one ``all_gather_into_tensor`` per column, enqueued on the current stream.  Without a process group (or world 1) the
    shard IS the column and is copied."""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    for k, src in shards.items():
        if multi:
            dist.all_gather_into_tensor(full[k], src)
        else:
            full[k].copy_(src, non_blocking=True)


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks when a process group is initialised (no-op otherwise)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def curve_stats(sum_, sumsq, count):
    """Mean and (population) standard deviation per episode from the reduced accumulators."""
    n = count.clamp(min=1)
    mean = sum_ / n
    var = (sumsq / n - mean * mean).clamp(min=0)
    return mean, var.sqrt()
