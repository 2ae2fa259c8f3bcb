"""Sentence sharding of a batch across the GPUs of one box.

Sentences are independent, so the DMV path and the alignment path need **no collective**: each rank parses its own
shard.  What this module provides is the host logic around that:

* ``shard_indices``     -- which sentences a rank takes.  Batches arrive sorted by length (descending, reference
  ``datamodule/sampler.py:135-136``); dealing them out round-robin keeps the cubic chart cost ``sum T(N_b)`` balanced.
* ``shard_batch``       -- slice merged score tensors / lengths for this rank.
* ``gather_heads``      -- bulk decode: collect every rank's predicted heads on all ranks in the original order
  (what ``pipeline.py:234-240`` does with ``all_gather_object`` on Python dicts, here on one int64 tensor).

Works with any ``torch.distributed`` backend (``nccl`` on the B200 box, ``gloo`` in the CPU tests).
"""
from __future__ import annotations

from typing import Optional

import torch


def chart_cost(lengths: torch.Tensor) -> torch.Tensor:
    """Split-point terms per chart sweep, T(N) = N^3 - N with N = len + 1."""
    n = lengths.to(torch.float64) + 1
    return n ** 3 - n


def shard_indices(num_sentences: int, rank: int, world_size: int) -> torch.Tensor:
    """Round-robin deal: rank r takes sentences r, r + W, r + 2W, ... (balanced for length-sorted batches)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    if rank >= num_sentences:
        return torch.empty(0, dtype=torch.int64)
    return torch.arange(rank, num_sentences, world_size)


def shard_batch(dec: torch.Tensor, attach: torch.Tensor, lengths: torch.Tensor, rank: int, world_size: int):
    """This rank's slice of a merged batch, plus the indices it came from."""
    idx = shard_indices(lengths.shape[0], rank, world_size).to(lengths.device)
    return dec.index_select(0, idx.to(dec.device)), attach.index_select(0, idx.to(attach.device)), lengths.index_select(0, idx), idx


def gather_heads(local_heads: torch.Tensor, num_sentences: int, rank: Optional[int] = None,
                 world_size: Optional[int] = None, group=None) -> torch.Tensor:
    """All ranks receive heads ``[num_sentences, N]`` in the original sentence order.

    ``local_heads`` is this rank's ``[ceil-or-floor(num_sentences / W), N]`` int64 tensor produced from ``shard_batch``.
    """
    import torch.distributed as dist

    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    N = local_heads.shape[1]
    out = local_heads.new_zeros((num_sentences, N))
    if world_size == 1:
        out[shard_indices(num_sentences, 0, 1).to(out.device)] = local_heads
        return out
    per = (num_sentences + world_size - 1) // world_size
    padded = local_heads.new_zeros((per, N))
    padded[: local_heads.shape[0]] = local_heads
    parts = [torch.empty_like(padded) for _ in range(world_size)]
    dist.all_gather(parts, padded, group=group)
    for r, part in enumerate(parts):
        idx = shard_indices(num_sentences, r, world_size).to(out.device)
        out[idx] = part[: idx.numel()]
    return out


def imbalance(lengths: torch.Tensor, world_size: int) -> float:
    """max / mean of the per-rank chart cost under ``shard_indices`` (1.0 = perfectly balanced)."""
    cost = chart_cost(lengths)
    per_rank = torch.stack([cost[shard_indices(len(lengths), r, world_size)].sum() for r in range(world_size)])
    return float(per_rank.max() / per_rank.mean())
