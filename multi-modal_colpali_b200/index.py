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
                      pack_queries, resolve_device)

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
                        payloads: Optional[Sequence[Any]] = None) -> np.ndarray:
        """Ingestion fusion: encoder hidden states ``[B, S, H]`` -> fused projection + L2-normalise +
        mask (K3) -> ragged page store, in HBM all the way (replaces ``embedding.tolist()`` + HTTP upsert,
        functions.py:843-865)."""
        from .head import project_normalize

        emb = project_normalize(hidden.to(self.device), weight, bias, attention_mask)
        return self.add_padded(emb, attention_mask.to(self.device), ids=ids, payloads=payloads)

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

    # -- persistence ------------------------------------------------------------------------------
    _CHUNK_ROWS = 1 << 22   # 1 GiB of 16-bit rows per copy

    def page_tables(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(offsets int64 [n+1], ids int64 [n], clamp uint8 [n]) copied to the host."""
        st = self._as_store()
        ids = _wrap_device(self._lib.lis_index_ids(self._h), (len(self),), torch.int64, self.device)
        return st.offsets.cpu().numpy(), ids.cpu().numpy(), st.clamp.cpu().numpy()

    def save(self, path) -> None:
        """Write the index to a directory: ``meta.json``, raw row planes (``tokens.bin`` [+ ``tokens_lo.bin``]),
        ``offsets.npy`` / ``ids.npy`` / ``clamp.npy`` and ``payloads.pkl``.  The row files are exactly the HBM
        layout, so loading is a straight copy (memory-mapped, chunked) with no re-encoding."""
        import json
        import pickle
        from pathlib import Path

        d = Path(path)
        d.mkdir(parents=True, exist_ok=True)
        n, rows = len(self), self.num_rows
        off, ids, clamp = self.page_tables() if n else (np.zeros(1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.uint8))
        np.save(d / "offsets.npy", off); np.save(d / "ids.npy", ids); np.save(d / "clamp.npy", clamp)
        planes = 2 if self.dtype == torch.float32 else 1
        with torch.cuda.device(self.device):
            for pl in range(planes):
                with open(d / ("tokens_lo.bin" if pl else "tokens.bin"), "wb") as f:
                    for r0 in range(0, rows, self._CHUNK_ROWS):
                        nr = min(self._CHUNK_ROWS, rows - r0)
                        buf = torch.empty((nr, N.DIM), dtype=torch.int16)
                        N.check(self._lib.lis_index_read_plane(self._h, pl, r0, nr, buf.data_ptr(), _stream(self.device)))
                        f.write(buf.numpy().tobytes())
        with open(d / "payloads.pkl", "wb") as f:
            pickle.dump(self.payloads, f)
        (d / "meta.json").write_text(json.dumps({
            "format": "lis-index-v1", "dtype": str(self.dtype).split(".")[-1], "n_pages": n, "n_rows": rows, "dim": N.DIM,
            "planes": planes}))

    @classmethod
    def load(cls, path, device=None, capacity_rows: Optional[int] = None, capacity_pages: Optional[int] = None
             ) -> "LateInteractionIndex":
        """Inverse of :meth:`save`; capacities default to the stored sizes (pass larger ones to keep adding)."""
        import json
        import pickle
        from pathlib import Path

        d = Path(path)
        meta = json.loads((d / "meta.json").read_text())
        if meta.get("format") != "lis-index-v1" or meta.get("dim") != N.DIM:
            raise ValueError(f"{d} is not a lis-index-v1 directory")
        dtype = getattr(torch, meta["dtype"])
        n, rows = int(meta["n_pages"]), int(meta["n_rows"])
        idx = cls(max(capacity_rows or rows, rows, 1), max(capacity_pages or n, n, 1), dtype=dtype, device=device)
        with torch.cuda.device(idx.device):
            for pl in range(int(meta["planes"])):
                if rows == 0:
                    break
                mm = np.memmap(d / ("tokens_lo.bin" if pl else "tokens.bin"), dtype=np.int16, mode="r", shape=(rows, N.DIM))
                for r0 in range(0, rows, cls._CHUNK_ROWS):
                    nr = min(cls._CHUNK_ROWS, rows - r0)
                    chunk = np.ascontiguousarray(mm[r0:r0 + nr])
                    N.check(idx._lib.lis_index_write_rows(idx._h, pl, r0, nr, chunk.ctypes.data, _stream(idx.device)))
            off = np.ascontiguousarray(np.load(d / "offsets.npy"), dtype=np.int64)
            ids = np.ascontiguousarray(np.load(d / "ids.npy"), dtype=np.int64)
            clamp = np.ascontiguousarray(np.load(d / "clamp.npy"), dtype=np.uint8)
            if len(off) != n + 1 or len(ids) != n or len(clamp) != n or (n and off[-1] != rows):
                raise ValueError(f"{d}: page tables are inconsistent with meta.json")
            N.check(idx._lib.lis_index_set_tables(idx._h, off.ctypes.data, ids.ctypes.data if n else None,
                                                  clamp.ctypes.data if n else None, n, _stream(idx.device)) if n else 0)
        with open(d / "payloads.pkl", "rb") as f:
            idx.payloads = pickle.load(f)
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
                                               plan.n_seg, plan.n_mtiles, seg_first, plan.nq, _ROUND[round_mode],
                                               int(k), out_s.data_ptr(), out_i.data_ptr(), _stream(self.device)))
        return out_s, out_i

    def search(self, qs: TensorOrList, k: int, round_mode: str = "f32") -> Tuple[torch.Tensor, torch.Tensor]:
        """Host-facing search: (scores fp32 [nq,k], page ids int64 [nq,k]) on the CPU, best first,
        ties broken by ascending id; slots beyond the corpus size hold (-inf, -1)."""
        s, i = self.search_device(qs, k, round_mode)
        return s.cpu(), i.cpu()

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
