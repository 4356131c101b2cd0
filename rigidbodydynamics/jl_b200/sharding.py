"""Batch sharding across the GPUs of one box: samples are independent, the flattened mechanism is
replicated in every process, so the data path has NO collective.  The only optional communication is the final gather of
v̇ (``gather_columns``), an NCCL all-gather over NVLink on GPUs (gloo on CPU in the tests).

One process per GPU, launched by ``torchrun``; each rank evaluates ``shard_bounds(B, world, rank)``.
"""
from __future__ import annotations

from typing import Tuple

import torch


def shard_bounds(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the batch owned by ``rank``: sizes differ by at most one, earlier ranks get the extra."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_columns(local: torch.Tensor, B: int, group=None) -> torch.Tensor:
    """All-gather ``[rows, n_local]`` shards (batch index fastest) into the full ``[rows, B]`` array on every rank.

    The result layout keeps the batch fastest, so the gather is done on the transposed ``[n_local, rows]`` view: shards are
    then contiguous slabs of the output and no post-gather transpose pass over the full array is needed on the receive side
    other than the final ``.t()`` view."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = local.shape[0]
    sizes = [shard_bounds(B, world, r) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    if local.shape[1] != hi - lo:
        raise ValueError("local shard has the wrong number of samples")
    send = torch.zeros((nmax, rows), dtype=local.dtype, device=local.device)
    send[: hi - lo] = local.t()
    recv = torch.empty((world * nmax, rows), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    out = torch.empty((B, rows), dtype=local.dtype, device=local.device)
    for r, (a, b) in enumerate(sizes):
        out[a:b] = recv[r * nmax: r * nmax + (b - a)]
    return out.t()
