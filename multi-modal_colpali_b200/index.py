"""GPU-resident page index + top-k search: the in-process replacement for the reference's two
search routes --

* ``score_results`` (05_experiment02.py:200-236): stack the whole corpus, ``score_multi_vector``,
  ``topk`` per query, gather page metadata;
* ``retrieve_colpali`` (functions.py:884-929): Qdrant ``query_points`` on a multivector MAX_SIM
  collection (schema 01_create_context_qdrant.py:208-222).

The index object is a thin handle on the C-side ``lis_index`` (tokens / offsets / ids live in HBM
and are owned by the library); payloads stay in a host-side dict keyed by page id.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native as N
from .scoring import (TensorOrList, _DTYPES, _ROUND, _as_list, _check_rows, _stream, clamp_flags,
                      pack_queries, plan_queries, resolve_device)

_TORCH_DTYPE = {N.LIS_BF16: torch.bfloat16, N.LIS_F16: torch.float16}


class LateInteractionIndex:
    """Ragged multi-vector page store on one GPU with fused MaxSim + top-k search.

    ``capacity_rows`` / ``capacity_pages`` are fixed at construction (HBM is sized once; a 180 GB
    B200 holds ~680 M token rows = 660 k ColPali pages of 1030 tokens)."""

    def __init__(self, capacity_rows: int, capacity_pages: int, dtype: torch.dtype = torch.bfloat16,
                 device: Union[str, torch.device, None] = None):
        if dtype not in _DTYPES and dtype != torch.float32:
            raise NotImplementedError(f"index dtype {dtype}: bfloat16, float16 or float32")
        self.device = resolve_device(device)
        self.dtype = dtype
        self._lib = N.load()
        self._h = C.c_void_p()
        code = N.LIS_F32X2 if dtype == torch.float32 else _DTYPES[dtype]   # fp32 -> two bf16 planes
        N.check(self._lib.lis_index_create(C.byref(self._h), self.device.index, code,
                                           int(capacity_rows), int(capacity_pages)))
        self.payloads: Dict[int, Any] = {}
        self._cap = (int(capacity_rows), int(capacity_pages))

    @property
    def capacity(self) -> Tuple[int, int]:
        """(token rows, pages) the HBM store can hold before :meth:`reserve` has to grow it."""
        return self._cap

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.lis_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self._lib.lis_index_num_pages(self._h))

    @property
    def num_rows(self) -> int:
        return int(self._lib.lis_index_num_rows(self._h))

    # -- ingestion --------------------------------------------------------------------------------
    def add(self, pages: TensorOrList, ids: Optional[Sequence[int]] = None,
            payloads: Optional[Sequence[Any]] = None, zero_pad_block: Optional[int] = None) -> np.ndarray:
        """Append pages (list of ``[n_tok,128]`` tensors or one ``[n, S, 128]`` tensor, host or device).

        ``zero_pad_block``: when set, pages shorter than the longest page of their block of that many
        consecutive pages get the reference's zero-padding semantics (per-token max clamped at 0), as
        ``score_multi_vector`` does with its 128-page batches.  Returns the assigned ids."""
        pl = _as_list(pages)
        if not pl:
            return np.zeros(0, np.int64)
        for t in pl:
            _check_rows(t, "page")
        lens = np.asarray([int(t.shape[0]) for t in pl], dtype=np.int32)
        n = len(pl)
        first = len(self)
        id_arr = np.arange(first, first + n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
        if id_arr.shape != (n,):
            raise ValueError("ids must have one entry per page")
        clamp = clamp_flags(lens, zero_pad_block) if zero_pad_block else None
        if isinstance(pages, torch.Tensor):
            flat = pages.reshape(-1, N.DIM)
        else:
            flat = torch.cat(pl, dim=0)
        flat = flat.to(self.dtype).contiguous()
        with torch.cuda.device(self.device):
            N.check(self._lib.lis_index_add(self._h, flat.data_ptr(), lens.ctypes.data, id_arr.ctypes.data,
                                            None if clamp is None else clamp.ctypes.data, n,
                                            _stream(self.device)))
        if payloads is not None:
            if len(payloads) != n:
                raise ValueError("payloads must have one entry per page")
            for i, pay in zip(id_arr.tolist(), payloads):
                self.payloads[i] = pay
        return id_arr

    def add_padded(self, embeddings: torch.Tensor, attention_mask: torch.Tensor, ids: Optional[Sequence[int]] = None,
                   payloads: Optional[Sequence[Any]] = None) -> np.ndarray:
        """Append a padded encoder batch ``[B, S, 128]`` (what ``model(**batch)`` returns: rows where
        ``attention_mask == 0`` are zero, on the left for ColQwen, on the right for ColPali) WITHOUT
        storing the pad rows: they are dropped and the page is flagged for the reference's zero-padding
        semantics instead (a zero row contributes similarity 0 to every max, i.e. ``max(v, 0)``), which is
        bit-identical to scoring the padded tensor and costs none of its bandwidth."""
        if embeddings.dim() != 3 or embeddings.shape[-1] != N.DIM or attention_mask.shape != embeddings.shape[:2]:
            raise ValueError("expected embeddings [B, S, 128] and attention_mask [B, S]")
        keep = attention_mask != 0
        lens = keep.sum(dim=1).to(torch.int32).cpu().numpy()
        rows = embeddings[keep.to(embeddings.device)]                 # ragged rows, page after page
        n, s_pad = embeddings.shape[0], embeddings.shape[1]
        first = len(self)
        id_arr = np.arange(first, first + n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
        if id_arr.shape != (n,):
            raise ValueError("ids must have one entry per page")
        clamp = (lens < s_pad).astype(np.uint8)
        flat = rows.to(self.dtype).contiguous()
        with torch.cuda.device(self.device):
            N.check(self._lib.lis_index_add(self._h, flat.data_ptr() if flat.numel() else None, lens.ctypes.data,
                                            id_arr.ctypes.data, clamp.ctypes.data, n, _stream(self.device)))
        if payloads is not None:
            if len(payloads) != n:
                raise ValueError("payloads must have one entry per page")
            for i, pay in zip(id_arr.tolist(), payloads):
                self.payloads[i] = pay
        return id_arr

    def add_from_hidden(self, hidden: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                        attention_mask: torch.Tensor, ids: Optional[Sequence[int]] = None,
                        payloads: Optional[Sequence[Any]] = None, round_mode: str = "reference") -> np.ndarray:
        """Ingestion fusion: encoder hidden states ``[B, S, H]`` -> K3 (projection + L2-normalise +
        mask) whose epilogue writes every kept row at ``store_row(page) + exclusive_prefix(mask)`` of the ragged
        page store.  No dense ``[B, S, 128]`` tensor, no gather, no copy: page lengths, clamp flags and destination
        rows are computed on the device (``lis_index_add_projected``); replaces ``model head -> tolist() -> HTTP
        upsert`` (functions.py:838-865)."""
        from .head import _ROUND as _HEAD_ROUND, mask_for_kernel

        if hidden.dim() != 3 or attention_mask.shape != hidden.shape[:2]:
            raise ValueError("expected hidden [B, S, H] and attention_mask [B, S]")
        if self.dtype not in _DTYPES:
            raise NotImplementedError("add_from_hidden: the encoder head is 16-bit; use a bf16/fp16 index")
        if round_mode not in _HEAD_ROUND:
            raise ValueError(f"round_mode must be one of {sorted(_HEAD_ROUND)}")
        n, s_pad, hdim = hidden.shape
        if weight.shape != (N.DIM, hdim):
            raise ValueError(f"weight must be [{N.DIM}, {hdim}], got {tuple(weight.shape)}")
        first = len(self)
        id_arr = np.arange(first, first + n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
        if id_arr.shape != (n,):
            raise ValueError("ids must have one entry per page")
        if payloads is not None and len(payloads) != n:
            raise ValueError("payloads must have one entry per page")
        h = hidden.to(device=self.device, dtype=self.dtype).contiguous()
        w = weight.to(device=self.device, dtype=self.dtype).contiguous()
        b = None if bias is None else bias.to(device=self.device, dtype=self.dtype).contiguous()
        m = mask_for_kernel(attention_mask, self.device)
        with torch.cuda.device(self.device):
            N.check(self._lib.lis_index_add_projected(self._h, h.data_ptr(), n, s_pad, hdim, w.data_ptr(),
                                                      None if b is None else b.data_ptr(), m.data_ptr(),
                                                      m.element_size(), _HEAD_ROUND[round_mode], id_arr.ctypes.data,
                                                      _stream(self.device)))
        if payloads is not None:
            for i, pay in zip(id_arr.tolist(), payloads):
                self.payloads[i] = pay
        return id_arr

    def page_lens(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        """Token rows of pages ``[first, first+n)`` (int32, host)."""
        n = len(self) - first if n is None else n
        out = np.zeros(max(n, 0), np.int32)
        if n > 0:
            with torch.cuda.device(self.device):
                N.check(self._lib.lis_index_page_lens(self._h, int(first), int(n), out.ctypes.data, _stream(self.device)))
        return out

    def fill_synthetic(self, n_pages: int, page_len: Union[int, Sequence[int]], seed: int, id_base: int = 0) -> None:
        """Append unit-norm pseudo-random pages generated on the device (benchmarks; see lis.h)."""
        lens = None
        fixed = 0
        if isinstance(page_len, (int, np.integer)):
            fixed = int(page_len)
        else:
            lens = np.ascontiguousarray(page_len, dtype=np.int32)
            if lens.shape != (n_pages,):
                raise ValueError("page_len must be an int or one length per page")
        with torch.cuda.device(self.device):
            N.check(self._lib.lis_index_fill_synthetic(self._h, int(n_pages), None if lens is None else lens.ctypes.data,
                                                       fixed, int(seed), int(id_base), _stream(self.device)))

    def read_rows(self, row0: int, n_rows: int) -> torch.Tensor:
        """Copy token rows back to the host (tests / debugging)."""
        out = torch.empty((n_rows, N.DIM), dtype=self.dtype)
        with torch.cuda.device(self.device):
            N.check(self._lib.lis_index_read_rows(self._h, int(row0), int(n_rows), out.data_ptr(), _stream(self.device)))
        return out

    # -- capacity / removal -------------------------------------------------------------------------
    def reserve(self, capacity_rows: int, capacity_pages: int) -> None:
        """Grow the HBM store to at least the given capacities (device-to-device copy of the row planes; the old
        store is released afterwards, so twice the current size is resident for a moment)."""
        rows, n = self.num_rows, len(self)
        cap_r, cap_p = max(int(capacity_rows), rows, 1), max(int(capacity_pages), n, 1)
        code = N.LIS_F32X2 if self.dtype == torch.float32 else _DTYPES[self.dtype]
        new_h = C.c_void_p()
        N.check(self._lib.lis_index_create(C.byref(new_h), self.device.index, code, cap_r, cap_p))
        try:
            with torch.cuda.device(self.device):
                st = _stream(self.device)
                if rows:
                    N.check(self._lib.lis_index_write_rows(new_h, 0, 0, rows, self._lib.lis_index_tokens(self._h), st))
                    if self.dtype == torch.float32:
                        N.check(self._lib.lis_index_write_rows(new_h, 1, 0, rows, self._lib.lis_index_tokens_lo(self._h), st))
                if n:
                    off, ids, clamp = self.page_tables()
                    N.check(self._lib.lis_index_set_tables(new_h, off.ctypes.data, ids.ctypes.data, clamp.ctypes.data, n, st))
        except BaseException:
            self._lib.lis_index_destroy(new_h)
            raise
        self._lib.lis_index_destroy(self._h)
        self._h = new_h
        self._cap = (cap_r, cap_p)

    def tombstone(self, page: int) -> None:
        """Hide page number ``page`` (position, not id) from all future searches (``lis_index_tombstone``)."""
        with torch.cuda.device(self.device):
            N.check(self._lib.lis_index_tombstone(self._h, int(page), _stream(self.device)))

    # -- persistence ------------------------------------------------------------------------------
    FORMAT = "lis-index-v2"

    def page_tables(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(offsets int64 [n+1], ids int64 [n], clamp uint8 [n]) copied to the host."""
        st = self._as_store()
        ids = _wrap_device(self._lib.lis_index_ids(self._h), (len(self),), torch.int64, self.device)
        return (np.ascontiguousarray(st.offsets.cpu().numpy()), np.ascontiguousarray(ids.cpu().numpy()),
                np.ascontiguousarray(st.clamp.cpu().numpy()))

    def _write_shard(self, d, p0: int, p1: int, off: np.ndarray, ids: np.ndarray, clamp: np.ndarray,
                     io_threads: int = 0) -> dict:
        """One shard directory: pages [p0, p1) of this index.  Row files are exactly the HBM layout."""
        import json

        d.mkdir(parents=True, exist_ok=True)
        r0, r1 = int(off[p0]), int(off[p1])
        np.save(d / "offsets.npy", off[p0:p1 + 1] - off[p0])
        np.save(d / "ids.npy", ids[p0:p1])
        np.save(d / "clamp.npy", clamp[p0:p1])
        planes = 2 if self.dtype == torch.float32 else 1
        with torch.cuda.device(self.device):
            for pl in range(planes):
                f = d / ("tokens_lo.bin" if pl else "tokens.bin")
                f.write_bytes(b"")
                N.check(self._lib.lis_index_save_rows(self._h, pl, r0, r1 - r0, str(f).encode(), 0, io_threads,
                                                      _stream(self.device)))
        pays = {int(i): self.payloads[int(i)] for i in ids[p0:p1].tolist() if int(i) in self.payloads}
        try:
            (d / "payloads.json").write_text(json.dumps({str(k): v for k, v in pays.items()}))
            payload_file = "payloads.json"
        except (TypeError, ValueError):
            import pickle

            with open(d / "payloads.pkl", "wb") as fh:
                pickle.dump(pays, fh)
            payload_file = "payloads.pkl"
        return {"dir": d.name, "n_pages": p1 - p0, "n_rows": r1 - r0, "payloads": payload_file,
                "id_min": int(ids[p0:p1].min()) if p1 > p0 else -1, "id_max": int(ids[p0:p1].max()) if p1 > p0 else -1}

    def save(self, path, shards: int = 1, io_threads: int = 0) -> dict:
        """Write the index as a sharded directory: ``manifest.json`` + ``shard-00000/ ...`` each holding the raw row
        planes (``tokens.bin`` [+ ``tokens_lo.bin``], exactly as they sit in HBM), ``offsets.npy`` / ``ids.npy`` /
        ``clamp.npy`` and the payloads (JSON when they are JSON-serialisable, else pickle).  ``shards``: how many
        pieces of roughly equal token count to cut this index into (a later multi-GPU load maps shards to ranks
        without touching the row files).  Rows leave the device through a pinned double buffer
        (``lis_index_save_rows``).  Returns the manifest."""
        import json
        from pathlib import Path

        from .sharded import balanced_shard_ranges

        d = Path(path)
        d.mkdir(parents=True, exist_ok=True)
        n = len(self)
        off, ids, clamp = self.page_tables() if n else (np.zeros(1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.uint8))
        ranges = balanced_shard_ranges(np.diff(off), max(1, min(int(shards), max(n, 1)))) if n else [(0, 0)]
        entries = [self._write_shard(d / f"shard-{i:05d}", a, b, off, ids, clamp, io_threads) for i, (a, b) in enumerate(ranges)]
        manifest = {"format": self.FORMAT, "dtype": str(self.dtype).split(".")[-1], "dim": N.DIM,
                    "planes": 2 if self.dtype == torch.float32 else 1, "n_pages": n, "n_rows": self.num_rows,
                    "row_bytes": 2 * N.DIM, "shards": entries}
        (d / "manifest.json").write_text(json.dumps(manifest, indent=1))
        return manifest

    @staticmethod
    def read_manifest(path) -> dict:
        import json
        from pathlib import Path

        d = Path(path)
        if (d / "manifest.json").exists():
            m = json.loads((d / "manifest.json").read_text())
            if m.get("format") != LateInteractionIndex.FORMAT or m.get("dim") != N.DIM:
                raise ValueError(f"{d} is not a {LateInteractionIndex.FORMAT} directory")
            return m
        if (d / "meta.json").exists():      # round-1 layout: one unsharded directory
            m = json.loads((d / "meta.json").read_text())
            if m.get("format") != "lis-index-v1" or m.get("dim") != N.DIM:
                raise ValueError(f"{d} is not a lis-index directory")
            return {"format": LateInteractionIndex.FORMAT, "dtype": m["dtype"], "dim": N.DIM, "planes": m["planes"],
                    "n_pages": m["n_pages"], "n_rows": m["n_rows"], "row_bytes": 2 * N.DIM,
                    "shards": [{"dir": ".", "n_pages": m["n_pages"], "n_rows": m["n_rows"], "payloads": "payloads.pkl"}]}
        raise ValueError(f"{d} holds no manifest.json")

    @classmethod
    def load(cls, path, device=None, capacity_rows: Optional[int] = None, capacity_pages: Optional[int] = None,
             shard_ids: Optional[Sequence[int]] = None, allow_pickle: bool = False, io_threads: int = 0
             ) -> "LateInteractionIndex":
        """Inverse of :meth:`save`.  ``shard_ids``: which shards of the directory to load, in order (default: all) --
        this is how a rank of a multi-GPU job picks its part (:meth:`ShardedIndex.load`).  Rows go from the files to
        HBM through a pinned double buffer with parallel reads (``lis_index_load_rows``); capacities default to the
        loaded sizes.  Pickled payloads are only read with ``allow_pickle=True`` (unpickling runs code: the directory
        must be trusted); JSON payloads need no flag."""
        import json
        from pathlib import Path

        d = Path(path)
        man = cls.read_manifest(d)
        dtype = getattr(torch, man["dtype"])
        shards = man["shards"]
        pick = list(range(len(shards))) if shard_ids is None else [int(i) for i in shard_ids]
        rows = sum(int(shards[i]["n_rows"]) for i in pick)
        n = sum(int(shards[i]["n_pages"]) for i in pick)
        idx = cls(max(capacity_rows or rows, rows, 1), max(capacity_pages or n, n, 1), dtype=dtype, device=device)
        offs, idl, cll = [np.zeros(1, np.int64)], [], []
        row0 = 0
        with torch.cuda.device(idx.device):
            for i in pick:
                sd = d / shards[i]["dir"]
                nr = int(shards[i]["n_rows"])
                for pl in range(int(man["planes"])):
                    f = sd / ("tokens_lo.bin" if pl else "tokens.bin")
                    if nr and f.stat().st_size != nr * 2 * N.DIM:
                        raise ValueError(f"{f}: size does not match the manifest")
                    N.check(idx._lib.lis_index_load_rows(idx._h, pl, row0, nr, str(f).encode(), 0, io_threads,
                                                         _stream(idx.device)))
                off = np.load(sd / "offsets.npy").astype(np.int64)
                if len(off) != int(shards[i]["n_pages"]) + 1 or off[0] != 0 or off[-1] != nr:
                    raise ValueError(f"{sd}: page tables are inconsistent with the manifest")
                offs.append(off[1:] + row0)
                idl.append(np.load(sd / "ids.npy").astype(np.int64))
                cll.append(np.load(sd / "clamp.npy").astype(np.uint8))
                row0 += nr
                pf = shards[i].get("payloads", "payloads.json")
                if pf.endswith(".json") and (sd / pf).exists():
                    idx.payloads.update({int(k): v for k, v in json.loads((sd / pf).read_text()).items()})
                elif (sd / pf).exists():
                    if not allow_pickle:
                        raise ValueError(f"{sd / pf} is a pickle; pass allow_pickle=True only for directories you trust")
                    import pickle

                    with open(sd / pf, "rb") as fh:
                        idx.payloads.update(pickle.load(fh))
            if n:
                off = np.ascontiguousarray(np.concatenate(offs))
                ids = np.ascontiguousarray(np.concatenate(idl))
                clamp = np.ascontiguousarray(np.concatenate(cll))
                N.check(idx._lib.lis_index_set_tables(idx._h, off.ctypes.data, ids.ctypes.data, clamp.ctypes.data, n,
                                                      _stream(idx.device)))
        return idx

    # -- search -----------------------------------------------------------------------------------
    def search_device(self, qs: TensorOrList, k: int, round_mode: str = "f32") -> Tuple[torch.Tensor, torch.Tensor]:
        """MaxSim + top-k entirely on the device; returns device tensors (scores fp32 [nq,k], ids int64 [nq,k])."""
        if len(qs) == 0:
            raise ValueError("No queries provided")
        if len(self) == 0:
            raise ValueError("No passages provided")
        if round_mode not in _ROUND:
            raise ValueError(f"round_mode must be one of {sorted(_ROUND)}")
        with torch.cuda.device(self.device):
            pq = pack_queries(qs, self.device, self.dtype)
            plan = pq.plan
            seg_lo, seg_hi, mt_seg, seg_first = pq.table_ptrs()
            out_s = torch.empty((plan.nq, k), dtype=torch.float32, device=self.device)
            out_i = torch.empty((plan.nq, k), dtype=torch.int64, device=self.device)
            N.check(self._lib.lis_index_search(self._h, pq.rows.data_ptr(),
                                               None if pq.rows_lo is None else pq.rows_lo.data_ptr(),
                                               pq.rows.shape[0], seg_lo, seg_hi, mt_seg,
                                               plan.n_seg, plan.n_mtiles, None if plan.direct else seg_first, plan.nq,
                                               _ROUND[round_mode],
                                               int(k), out_s.data_ptr(), out_i.data_ptr(), _stream(self.device)))
        return out_s, out_i

    def search(self, qs: TensorOrList, k: int, round_mode: str = "f32", comm=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Host-facing search: (scores fp32 [nq,k], page ids int64 [nq,k]) on the CPU, best first,
        ties broken by ascending id; slots beyond the corpus size hold (-inf, -1).

        One C call (``lis_index_search_sharded``): queries are packed on the host, and the whole device sequence
        -- upload, K1, segment sums, K2, [all-gather + merge when ``comm`` spans several ranks,] download -- replays
        as one CUDA graph per query shape.  ``comm``: a ``lis_comm`` handle (see :class:`ShardedIndex`)."""
        if len(qs) == 0:
            raise ValueError("No queries provided")
        if comm is None and len(self) == 0:
            raise ValueError("No passages provided")
        if round_mode not in _ROUND:
            raise ValueError(f"round_mode must be one of {sorted(_ROUND)}")
        k = int(k)
        rows, lens = _flatten_queries(qs, self.dtype)
        plan = plan_queries(lens)
        if plan.n_seg == 0:
            raise ValueError("No queries provided")
        out_s = torch.empty(plan.nq, k, dtype=torch.float32)
        out_i = torch.empty(plan.nq, k, dtype=torch.int64)
        # (the C call selects the index's device itself; no torch device guard on this latency path)
        seg_lo, seg_hi, mt_seg, seg_first = plan.host_ptrs
        N.check(self._lib.lis_index_search_sharded(
            self._h, comm, rows.data_ptr(), plan.n_rows, seg_lo, seg_hi, mt_seg, plan.n_seg, plan.n_mtiles, seg_first,
            plan.nq, _ROUND[round_mode], k, out_s.data_ptr(), out_i.data_ptr(),
            torch.cuda.current_stream(rows.device).cuda_stream if rows.is_cuda else None))
        return out_s, out_i

    def graph_stats(self) -> Tuple[int, int, int]:
        """(cached search graphs, captures, replays) of the one-shot search path."""
        cap, rep = C.c_int64(0), C.c_int64(0)
        n = self._lib.lis_index_graph_stats(self._h, C.byref(cap), C.byref(rep))
        return int(n), int(cap.value), int(rep.value)

    def scores(self, qs: TensorOrList, round_mode: str = "f32") -> torch.Tensor:
        """Full device fp32 ``[nq, n_pages]`` score matrix against the resident corpus."""
        from .scoring import PageStore, maxsim_scores_device

        with torch.cuda.device(self.device):
            pq = pack_queries(qs, self.device, self.dtype)
            store = self._as_store()
            return maxsim_scores_device(pq, store, round_mode)

    def _as_store(self):
        from .scoring import PageStore

        n, rows = len(self), self.num_rows
        plane = torch.bfloat16 if self.dtype == torch.float32 else self.dtype
        tok = _wrap_device(self._lib.lis_index_tokens(self._h), (rows, N.DIM), plane, self.device)
        off = _wrap_device(self._lib.lis_index_offsets(self._h), (n + 1,), torch.int64, self.device)
        cl = _wrap_device(self._lib.lis_index_clamp(self._h), (n,), torch.uint8, self.device)
        lo = None
        if self.dtype == torch.float32:
            lo = _wrap_device(self._lib.lis_index_tokens_lo(self._h), (rows, N.DIM), plane, self.device)
        return PageStore(tok, off, cl, n, lo, self.dtype)


def _flatten_queries(qs: TensorOrList, dtype: torch.dtype) -> Tuple[torch.Tensor, Tuple[int, ...]]:
    """Queries as ONE contiguous ``[rows, 128]`` matrix of the index dtype, where they already live (host
    tensors stay on the host: the C call uploads them together with the segment tables), plus their lengths.
    This sits on the latency path of a small-corpus search, hence the early exits for data that is already in shape."""
    if isinstance(qs, torch.Tensor) and qs.dim() == 3:
        nq, n_tok, d = qs.shape
        if d != N.DIM:
            raise ValueError(f"queries: embedding width {d} != {N.DIM}")
        lens = (n_tok,) * nq
        if qs.dtype == dtype and qs.is_contiguous():
            return qs.detach().view(nq * n_tok, d), lens
        flat = qs.reshape(nq * n_tok, d)
    else:
        ql = qs if isinstance(qs, (list, tuple)) else _as_list(qs)
        if len(ql) == 0:
            raise ValueError("No queries provided")
        dev, same_dev = ql[0].device, True
        for t in ql:
            _check_rows(t, "query")
            same_dev = same_dev and t.device == dev
        lens = tuple(t.shape[0] for t in ql)
        if not same_dev:
            ql = [t.cpu() for t in ql]
        flat = ql[0] if len(ql) == 1 else torch.cat(ql, dim=0)
        if flat.dtype == dtype and flat.is_contiguous():
            return flat.detach(), lens
    return flat.detach().to(dtype).contiguous(), lens


class _CudaArrayView:
    """Minimal ``__cuda_array_interface__`` carrier so torch can alias library-owned HBM."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], typestr: str):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}


def _wrap_device(ptr: int, shape: Tuple[int, ...], dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    if dtype in (torch.bfloat16, torch.float16):
        t = torch.as_tensor(_CudaArrayView(ptr, shape, "<i2"), device=device)
        return t.view(dtype)
    typestr = {torch.int64: "<i8", torch.uint8: "|u1", torch.float32: "<f4"}[dtype]
    return torch.as_tensor(_CudaArrayView(ptr, shape, typestr), device=device)


def topk_device(scores: torch.Tensor, k: int, ids: Optional[torch.Tensor] = None, id_base: int = 0
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """K2 on a device fp32 ``[nq, np]`` matrix (replaces ``query_scores.topk(top_k)``,
    05_experiment02.py:219).  ``ids`` int64 ``[np]`` maps columns to page ids; entries < 0 are
    excluded (this is how payload filters are applied).  Order: score desc, id asc."""
    lib = N.load()
    if scores.dim() != 2 or scores.dtype != torch.float32 or not scores.is_cuda:
        raise ValueError("scores must be a CUDA float32 [nq, np] tensor")
    if scores.stride(1) != 1:
        scores = scores.contiguous()
    nq, npg = scores.shape
    dev = scores.device
    with torch.cuda.device(dev):
        ws_bytes = int(lib.lis_topk_workspace_bytes(nq, npg, int(k)))
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        N.check(lib.lis_topk(scores.data_ptr(), scores.stride(0), nq, npg, None if ids is None else ids.data_ptr(),
                             int(id_base), int(k), out_s.data_ptr(), out_i.data_ptr(), ws.data_ptr(), ws.numel(),
                             _stream(dev)))
    return out_s, out_i


def merge_topk_device(cand_scores: torch.Tensor, cand_ids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge candidate lists ``[nq, n_cand]`` (e.g. allgathered per-GPU top-k) into the global top-k."""
    lib = N.load()
    if cand_scores.shape != cand_ids.shape or cand_scores.dim() != 2:
        raise ValueError("candidate scores / ids must both be [nq, n_cand]")
    cand_scores = cand_scores.contiguous()
    cand_ids = cand_ids.contiguous()
    nq, nc = cand_scores.shape
    dev = cand_scores.device
    with torch.cuda.device(dev):
        ws_bytes = int(lib.lis_topk_workspace_bytes(nq, nc, int(k)))
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        N.check(lib.lis_merge_topk(cand_scores.data_ptr(), cand_ids.data_ptr(), nq, nc, int(k), out_s.data_ptr(),
                                   out_i.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
    return out_s, out_i
