"""Batch-index sharding of one pass over the ranks of a box.

Samples are independent in eval mode (BatchNorm uses running stats, pro_b_gan_infer.py:106-107), so the pass
shards by contiguous batch-index ranges: rank r of g owns rows [lo, hi) of the [B, ...] index / latent
tensors; weights and the embedding tables are replicated.  The only exchange step is the all-gather that
reassembles the outputs (north_star) -- G's [B, E] predictions and D's [B] logits / probabilities.

Pure host arithmetic + torch.distributed plumbing: works on gloo (CPU tests) and nccl (B200).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous split; the first ``batch % world_size`` ranks get one extra row (ragged batches)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} of {world_size}")
    if batch < 0:
        raise ValueError("negative batch")
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(batch: int, world_size: int) -> list[int]:
    return [hi - lo for lo, hi in (shard_bounds(batch, world_size, r) for r in range(world_size))]


def take_shard(t: torch.Tensor, world_size: int, rank: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], world_size, rank)
    return t[lo:hi]


def all_gather_rows(local: torch.Tensor, batch: int, group=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """Reassemble a row-sharded tensor: every rank ends with the full [batch, ...] tensor.

    Even shards use one ``all_gather_into_tensor`` (NCCL: one ring/NVLS collective over NVLink); ragged shards
    fall back to the list form."""
    ws = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(batch, ws)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, its shard of {batch} is {sizes[rank]}")
    if out is None:
        out = torch.empty((batch,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:
        pieces = list(out.split(sizes, dim=0))
        if local.device.type == "cuda":
            dist.all_gather(pieces, local.contiguous(), group=group)
        else:  # gloo needs equal sizes: pad to the largest shard
            m = max(sizes)
            padded = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype)
            padded[: local.shape[0]] = local
            bufs = [torch.empty_like(padded) for _ in range(ws)]
            dist.all_gather(bufs, padded, group=group)
            for p, b, n in zip(pieces, bufs, sizes):
                p.copy_(b[:n])
    return out
