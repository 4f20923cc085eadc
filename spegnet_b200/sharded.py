"""Batch-sharded inference across the GPUs of one box.

Images are independent in eval mode, so the forward needs no data-path collective: image i goes to rank
``i % world`` with a full weight replica per GPU.  The only exchange is one ``all_gather`` per dataset of
small per-image rows (index + scores / integer mask statistics); dataset means are then taken in index order
on every rank, so the result is bit-identical for any world size (the reference averages plain per-image
scores: engine/evaluator.py:447-457, utils/metrics.py:268-275).

Everything here works on CPU tensors with the ``gloo`` backend too (tests/test_sharded_cpu.py).
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def world_info() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin shard: item i -> rank i % world (ranks differ by at most one item)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_items, world))


def padded_shard_size(n_items: int, world: int) -> int:
    return (n_items + world - 1) // world


def gather_rows(local_index: Sequence[int], local_rows: torch.Tensor, n_items: int) -> torch.Tensor:
    """All-gather per-item rows.  `local_rows` is [len(local_index), k]; returns [n_items, k] in item order on
    every rank.  Shards are padded to ceil(n/world) rows with index -1 so that ragged splits work."""
    rank, world = world_info()
    k = local_rows.shape[1]
    pad = padded_shard_size(n_items, world)
    if len(local_index) > pad or local_rows.shape[0] != len(local_index):
        raise ValueError("local rows do not match the shard")
    buf = torch.zeros(pad, k + 1, dtype=torch.float64, device=local_rows.device)
    buf[:, 0] = -1
    if len(local_index):
        buf[: len(local_index), 0] = torch.as_tensor(list(local_index), dtype=torch.float64, device=buf.device)
        buf[: len(local_index), 1:] = local_rows.to(torch.float64)
    if world > 1:
        out = torch.empty(world * pad, k + 1, dtype=torch.float64, device=buf.device)
        dist.all_gather_into_tensor(out, buf)
    else:
        out = buf
    valid = out[out[:, 0] >= 0]
    if valid.shape[0] != n_items:
        raise RuntimeError(f"gathered {valid.shape[0]} rows for {n_items} items (duplicate or missing shards)")
    order = torch.argsort(valid[:, 0])
    rows = valid[order]
    if not torch.equal(rows[:, 0].long(), torch.arange(n_items, device=rows.device)):
        raise RuntimeError("gathered item indices are not a permutation of 0..n-1")
    return rows[:, 1:]


def sharded_map(n_items: int, batch_size: int, fn: Callable[[List[int]], torch.Tensor]) -> torch.Tensor:
    """Run `fn(indices) -> [len(indices), k]` over this rank's shard in batches and gather all rows.
    `fn` is the per-batch work (forward + per-image statistics); returns [n_items, k] on every rank."""
    rank, world = world_info()
    mine = shard_indices(n_items, rank, world)
    chunks: List[torch.Tensor] = []
    for s in range(0, len(mine), batch_size):
        chunks.append(fn(mine[s:s + batch_size]))
    if chunks:
        local = torch.cat(chunks, dim=0)
    else:
        probe = fn([])  # an empty shard still has to agree on k
        local = probe
    return gather_rows(mine, local, n_items)


def mean_in_index_order(rows: torch.Tensor) -> torch.Tensor:
    """Plain mean over items, accumulated in index order in fp64 (independent of the world size)."""
    return rows.to(torch.float64).sum(dim=0) / rows.shape[0]
