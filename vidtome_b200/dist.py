"""Chunk-per-GPU sharding and the global-token exchange.

One process per GPU, `torch.distributed` for the plumbing (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  Two cases:

* `merge_global=False` — frame chunks of one denoising step are independent (generate.py:216-219): each rank
  takes its chunks (`shard_chunks`), weights are replicated, NOTHING crosses GPUs in the data path.  Every rank must
  hold the same generator state so that the per-block random target frames agree (`seed_all_ranks`).

* `merge_global=True` — in the reference the global token set is a sequential recurrence over chunks
  (patch.py:59-82; chunk k's block l sees what chunk k-1's block l left in `module.global_tokens`).  Sharded one
  chunk per GPU, the exchange is ONE all-gather of the local merged tokens per merged block
  (`exchange_global_tokens`): rank k then matches against the local tokens of rank (k-1) mod G.  This is the
  north-star variant, a semantic variant of the recurrence (it reproduces one step of it with "previous chunk's
  local tokens" in place of the running set); parity is defined against the reference's own `_2s` matcher fed
  the same global tokens (the attribute is public, patch.py:60).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_chunks(chunks: Sequence, rank_: Optional[int] = None, world_: Optional[int] = None) -> List:
    """Round-robin assignment of a step's chunk list to ranks (chunk i -> rank i % world)."""
    r = rank() if rank_ is None else rank_
    w = world() if world_ is None else world_
    return [c for i, c in enumerate(chunks) if i % w == r]


def seed_all_ranks(seed: int) -> None:
    """Same default-RNG state on every rank: the per-block generators are forked from it at first forward
    (patch.py:219-226), so all ranks then draw identical random target frames."""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def all_gather_tokens(local: torch.Tensor, group=None) -> List[torch.Tensor]:
    """All-gather of `[B, L_k, C]` token sets whose length L_k may differ per rank (ragged first chunk,
    generate.py:176-178).  One small all-gather of the lengths, one padded all-gather of the tokens."""
    w = world()
    if w == 1:
        return [local]
    B, L, C = local.shape
    lens = torch.tensor([L], device=local.device, dtype=torch.int64)
    all_lens = [torch.zeros_like(lens) for _ in range(w)]
    dist.all_gather(all_lens, lens, group=group)
    all_lens = [int(t.item()) for t in all_lens]
    Lmax = max(all_lens)
    padded = local if L == Lmax else torch.cat([local, local.new_zeros((B, Lmax - L, C))], dim=1)
    out = torch.empty((w * B, Lmax, C), dtype=local.dtype, device=local.device)   # rank-major concatenation
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    out = out.view(w, B, Lmax, C)
    return [out[k, :, : all_lens[k]] for k in range(w)]


def exchange_global_tokens(local_tokens: torch.Tensor, group=None) -> Optional[torch.Tensor]:
    """The single collective of the global-merge path: all-gather the local merged tokens of this block and
    return the set this rank matches against — the tokens of rank (k-1) mod G (None when running alone)."""
    w = world()
    if w == 1:
        return None
    gathered = all_gather_tokens(local_tokens, group=group)
    return gathered[(rank() - 1) % w].contiguous()
