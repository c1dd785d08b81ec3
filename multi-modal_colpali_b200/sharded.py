"""Corpus-sharded search: every rank (one process per GPU) owns a contiguous slice of the pages,
scores it and keeps a local top-k; candidates are exchanged with ONE all-gather over NCCL/NVLink and
merged by the same tournament kernel on every rank.  The reference has no multi-GPU path
(SURVEY.md section 2c); pages are independent, so this is the only exchange step.

The exchange itself lives behind the C-ABI (``lis_comm_init`` / ``lis_index_search_sharded``: K2's last pass
writes into the send buffer, ``ncclAllGather``, merge, one download, all inside one CUDA graph); this module
only creates the communicator (the NCCL id travels over ``torch.distributed``) and holds the shard.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _native as N
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
    """All-gather per-rank ``[nq, k]`` candidates into ``[nq, world*k]`` (rank-major columns) through
    ``torch.distributed`` -- the host-logic twin of the C path, used by the gloo tests and by callers that keep
    their candidates in torch tensors.  One collective: scores are bit-cast into the int64 payload next to the ids."""
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


class Communicator:
    """``lis_comm`` handle: one NCCL communicator per process, created from an id that rank 0 makes and
    ``torch.distributed`` (any backend) broadcasts."""

    def __init__(self, device: torch.device, group=None):
        self._lib = N.load()
        self._h = C.c_void_p()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        uid = (C.c_uint8 * N.COMM_ID_BYTES)()
        if self.world > 1:
            box = [None]
            if self.rank == 0:
                N.check(self._lib.lis_comm_unique_id(uid, N.COMM_ID_BYTES))
                box[0] = bytes(uid)
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            C.memmove(uid, box[0], N.COMM_ID_BYTES)
        with torch.cuda.device(device):
            N.check(self._lib.lis_comm_init(C.byref(self._h), uid, self.rank, self.world, device.index))

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.lis_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass


class ShardedIndex:
    """A :class:`LateInteractionIndex` per rank + the all-gather/merge step.

    Production: ``ShardedIndex(local_index)`` inside an initialised ``torch.distributed`` job -- ``search`` is one
    call into ``lis_index_search_sharded``.  Page ids must be global (``fill_synthetic(..., id_base=...)`` /
    ``add(..., ids=...)``).  An empty local shard is fine: it contributes padding and still joins the collective.

    ``local_search`` / ``merge`` are injectable so the host-side plumbing is testable on CPU with the
    gloo backend (tests/test_sharded_gloo.py); with them the exchange goes through ``gather_candidates``."""

    def __init__(self, local: Optional[LateInteractionIndex], group=None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None):
        self.local = local
        self.group = group
        self._injected = local_search is not None
        self._local_search = local_search or self._search_local_shard
        self._merge = merge or merge_topk_device
        self.comm: Optional[Communicator] = None
        if not self._injected and local is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
            self.comm = Communicator(local.device, group)

    @property
    def dtype(self):
        return getattr(self.local, "dtype", None)

    def _search_local_shard(self, qs, k: int, round_mode: str):
        if len(self.local) == 0:    # an empty shard contributes padding, which the merge ignores
            nq = len(qs)
            dev = self.local.device
            return (torch.full((nq, k), float("-inf"), dtype=torch.float32, device=dev),
                    torch.full((nq, k), -1, dtype=torch.int64, device=dev))
        return self.local.search_device(qs, k, round_mode)

    def search_device(self, qs, k: int, round_mode: str = "f32") -> Tuple[torch.Tensor, torch.Tensor]:
        """Torch-level path (device tensors in and out): local top-k, ``gather_candidates``, merge.  A local
        failure is reported to every rank BEFORE the collective, so no rank is left waiting in it."""
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        err: Optional[BaseException] = None
        s = i = None
        try:
            s, i = self._local_search(qs, k, round_mode)
        except BaseException as exc:   # noqa: BLE001 - re-raised below, after the ranks agreed
            if not multi:
                raise
            err = exc
        if not multi:
            return s, i
        dev = s.device if s is not None else (self.local.device if self.local is not None else torch.device("cpu"))
        flag = torch.tensor([1 if err is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        if int(flag.item()):
            if err is not None:
                raise err
            raise RuntimeError("sharded search failed on another rank")
        all_s, all_i = gather_candidates(s, i, self.group)
        return self._merge(all_s, all_i, k)

    def search(self, qs, k: int, round_mode: str = "f32") -> Tuple[torch.Tensor, torch.Tensor]:
        if self._injected or self.local is None:
            s, i = self.search_device(qs, k, round_mode)
            return s.cpu(), i.cpu()
        return self.local.search(qs, k, round_mode, comm=None if self.comm is None else self.comm.handle)

    def close(self) -> None:
        if self.comm is not None:
            self.comm.close()
            self.comm = None

    # -- persistence: one shard directory per rank + one manifest ---------------------------------------------
    def save(self, path, io_threads: int = 0) -> Optional[dict]:
        """Every rank writes its shard (``shard-<rank>/``, raw HBM row planes + page tables) concurrently; rank 0 then
        writes ``manifest.json`` listing the shards in rank order with their page / row counts (the record of how the
        corpus was cut: ``shard_range`` / ``balanced_shard_ranges``).  Same format as
        :meth:`LateInteractionIndex.save`, so the directory also loads on one GPU or on a different world size."""
        import json
        from pathlib import Path

        import numpy as np

        d = Path(path)
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        rank = dist.get_rank(self.group) if multi else 0
        world = dist.get_world_size(self.group) if multi else 1
        ix = self.local
        n = len(ix)
        off, ids, clamp = ix.page_tables() if n else (np.zeros(1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.uint8))
        d.mkdir(parents=True, exist_ok=True)
        entry = ix._write_shard(d / f"shard-{rank:05d}", 0, n, off, ids, clamp, io_threads)
        entries = [entry]
        if multi:
            entries = [None] * world
            dist.all_gather_object(entries, entry, group=self.group)
        if rank != 0:
            return None
        manifest = {"format": ix.FORMAT, "dtype": str(ix.dtype).split(".")[-1], "dim": N.DIM,
                    "planes": 2 if ix.dtype == torch.float32 else 1, "n_pages": sum(e["n_pages"] for e in entries),
                    "n_rows": sum(e["n_rows"] for e in entries), "row_bytes": 2 * N.DIM, "shards": entries,
                    "written_by_world": world}
        (d / "manifest.json").write_text(json.dumps(manifest, indent=1))
        return manifest

    @classmethod
    def load(cls, path, device=None, group=None, allow_pickle: bool = False, io_threads: int = 0) -> "ShardedIndex":
        """Each rank loads a contiguous run of the directory's shards, balanced by token rows (``assign_shards``), so
        a directory written by W ranks (or cut into S pieces by ``LateInteractionIndex.save(shards=S)``) loads on any
        world size; with S == world every rank reads exactly one shard.  Page ids are the stored (global) ones."""
        multi = dist.is_initialized() and dist.get_world_size(group) > 1
        rank = dist.get_rank(group) if multi else 0
        world = dist.get_world_size(group) if multi else 1
        man = LateInteractionIndex.read_manifest(path)
        mine = assign_shards([int(s["n_rows"]) for s in man["shards"]], world)[rank]
        local = LateInteractionIndex.load(path, device=device, shard_ids=list(range(*mine)), allow_pickle=allow_pickle,
                                          io_threads=io_threads)
        return cls(local, group=group)


def assign_shards(shard_rows: Sequence[int], world: int) -> list:
    """Contiguous shard ranges [begin, end) per rank, balanced by rows; one shard per rank when the counts match."""
    if len(shard_rows) == world:
        return [(r, r + 1) for r in range(world)]
    return balanced_shard_ranges(shard_rows, world)
