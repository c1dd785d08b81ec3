"""Corpus-sharded search: every rank (one process per GPU) owns a contiguous slice of the pages,
scores it and keeps a local top-k; candidates are exchanged with ONE all-gather over NCCL/NVLink and
merged by the same tournament kernel on every rank.  The reference has no multi-GPU path
(SURVEY.md section 2c); pages are independent, so this is the only exchange step.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .index import LateInteractionIndex, merge_topk_device


def shard_range(n_pages: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous page range [begin, end) of ``rank`` (sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world) or n_pages < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_pages, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def balanced_shard_ranges(page_lens: Sequence[int], world: int) -> list:
    """Contiguous ranges balanced by TOKEN count (ragged corpora: SURVEY.md section 8e)."""
    import numpy as np

    lens = np.asarray(page_lens, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(lens)])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        c = int(np.searchsorted(csum, target, side="left"))
        if c > 0 and (c >= len(csum) or abs(csum[c - 1] - target) <= abs(csum[c] - target)):
            c -= 1  # the boundary nearest to the ideal cut
        cuts.append(c)
    cuts.append(len(lens))
    cuts = [min(max(c, cuts[i - 1] if i else 0), len(lens)) for i, c in enumerate(cuts)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_candidates(scores: torch.Tensor, ids: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather per-rank ``[nq, k]`` candidates into ``[nq, world*k]`` (rank-major columns).
    One collective: scores are bit-cast into the int64 payload next to the ids."""
    world = dist.get_world_size(group)
    nq, k = scores.shape
    packed = torch.empty((nq, 2, k), dtype=torch.int64, device=scores.device)
    packed[:, 0, :] = scores.contiguous().view(torch.int32).to(torch.int64)
    packed[:, 1, :] = ids
    flat = torch.empty((world * nq, 2, k), dtype=torch.int64, device=scores.device)
    dist.all_gather_into_tensor(flat, packed, group=group)
    out = flat.view(world, nq, 2, k)
    all_s = out[:, :, 0, :].to(torch.int32).view(torch.float32).permute(1, 0, 2).reshape(nq, world * k)
    all_i = out[:, :, 1, :].permute(1, 0, 2).reshape(nq, world * k)
    return all_s.contiguous(), all_i.contiguous()


class ShardedIndex:
    """A :class:`LateInteractionIndex` per rank + the all-gather/merge step.

    ``local_search`` / ``merge`` are injectable so the host-side plumbing is testable on CPU with the
    gloo backend (tests/test_sharded_gloo.py); in production they are the CUDA kernels."""

    def __init__(self, local: Optional[LateInteractionIndex], group=None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None):
        self.local = local
        self.group = group
        self._local_search = local_search or (lambda qs, k, rm: local.search_device(qs, k, rm))
        self._merge = merge or merge_topk_device

    def search_device(self, qs, k: int, round_mode: str = "f32") -> Tuple[torch.Tensor, torch.Tensor]:
        s, i = self._local_search(qs, k, round_mode)
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return s, i
        all_s, all_i = gather_candidates(s, i, self.group)
        return self._merge(all_s, all_i, k)

    def search(self, qs, k: int, round_mode: str = "f32") -> Tuple[torch.Tensor, torch.Tensor]:
        s, i = self.search_device(qs, k, round_mode)
        return s.cpu(), i.cpu()
