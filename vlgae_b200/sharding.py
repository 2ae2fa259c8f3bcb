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
    if world_size == 1:
        return local_heads[:num_sentences].clone()
    # Heads are positions < N: they travel as int16 (4x fewer bytes than the int64 the API returns), in ONE collective
    # into a [W, per, N] buffer.  The round-robin deal puts sentence k W + r at (r, k), so the original order is a
    # transpose of that buffer -- one strided copy, fused with the widening back to int64.
    per = (num_sentences + world_size - 1) // world_size
    wire = torch.int16 if N <= 32767 else local_heads.dtype
    if local_heads.shape[0] == per:
        padded = local_heads.to(wire)  # one narrowing kernel
    else:                              # the last ranks hold one sentence less: pad
        padded = torch.zeros((per, N), dtype=wire, device=local_heads.device)
        padded[: local_heads.shape[0]] = local_heads
    parts = torch.empty((world_size, per, N), dtype=wire, device=local_heads.device)
    # (on the wire as raw bytes: neither NCCL nor gloo has an int16 type, and a gather needs none)
    src, dst = padded.view(torch.uint8), parts.view(torch.uint8)
    try:
        dist.all_gather_into_tensor(dst.view(world_size * per, -1), src, group=group)
    except (RuntimeError, NotImplementedError):  # a backend without the tensor form
        dist.all_gather(list(dst.unbind(0)), src, group=group)
    out = torch.empty((per, world_size, N), dtype=local_heads.dtype, device=local_heads.device)
    out.copy_(parts.permute(1, 0, 2))  # transpose + widen in one kernel
    return out.view(per * world_size, N)[:num_sentences]


def imbalance(lengths: torch.Tensor, world_size: int) -> float:
    """max / mean of the per-rank chart cost under ``shard_indices`` (1.0 = perfectly balanced)."""
    cost = chart_cost(lengths)
    per_rank = torch.stack([cost[shard_indices(len(lengths), r, world_size)].sum() for r in range(world_size)])
    return float(per_rank.max() / per_rank.mean())
