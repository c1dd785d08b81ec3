"""Drop-in for ``processor.score_multi_vector(qs, ps)`` (reference call site
``05_experiment02.py:214``; body in colpali-engine 0.3.13, arithmetic identical to HF
``processing_colpali.py:350-364``), backed by the fused sm_100a kernel.

PyTorch is used for device memory and streams only; all arithmetic happens in ``liblis.so``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from functools import cached_property, lru_cache
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native as N

TensorOrList = Union[torch.Tensor, Sequence[torch.Tensor]]

_DTYPES = {torch.bfloat16: N.LIS_BF16, torch.float16: N.LIS_F16}
_ROUND = {"f32": N.ROUND_F32, "reference": N.ROUND_REFERENCE}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def resolve_device(device: Union[str, torch.device, None]) -> torch.device:
    """The reference picks cuda -> mps -> cpu automatically; this engine is CUDA-only by design."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("multi-modal_colpali_b200 needs an sm_100 (B200) GPU; there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"device {dev} is not supported: the scoring engine is sm_100a CUDA only")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


# ------------------------------------------------------------------------------------------------
# Query packing (host logic; runs without a GPU)
# ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class QueryPlan:
    """Segmentation of back-to-back query rows at 128-row M-tile boundaries (``lis_plan_queries``)."""
    nq: int
    n_rows: int             # total query-token rows
    n_seg: int
    n_mtiles: int
    seg_query: np.ndarray   # int32 [n_seg]
    seg_lo: np.ndarray      # int32 [n_seg]
    seg_hi: np.ndarray      # int32 [n_seg]
    mt_seg: np.ndarray      # int32 [n_mtiles+1]
    seg_first: np.ndarray   # int32 [nq+1]  segments of query q = seg_first[q]..seg_first[q+1]

    @cached_property
    def direct(self) -> bool:
        """True when K1's output rows are the per-query scores: segment s belongs to query s, i.e. no query was
        cut and none is empty (an empty query owns no segment, so the counts alone cannot tell: lens [0, 100] give
        two segments of query 1)."""
        return self.n_seg == self.nq and bool((self.seg_query == np.arange(self.nq, dtype=np.int32)).all())

    @cached_property
    def host_ptrs(self) -> Tuple[int, int, int, Optional[int]]:
        """Host addresses of (seg_lo, seg_hi, mt_seg, seg_first or None when ``direct``) -- what the one-shot search
        passes to the C call.  Plans are cached and immutable, so the addresses are computed once per query shape."""
        return (self.seg_lo.ctypes.data, self.seg_hi.ctypes.data, self.mt_seg.ctypes.data,
                None if self.direct else self.seg_first.ctypes.data)


def plan_queries(q_lens: Sequence[int]) -> QueryPlan:
    if type(q_lens) is not tuple:
        q_lens = tuple(int(x) for x in q_lens)
    return _plan_cached(q_lens)


@lru_cache(maxsize=256)
def _plan_cached(q_lens: Tuple[int, ...]) -> QueryPlan:
    lib = N.load()
    nq = len(q_lens)
    lens = np.asarray(q_lens, dtype=np.int32)
    n_mt = C.c_int64(0)
    n_seg = lib.lis_plan_queries(lens.ctypes.data, nq, 0, None, None, None, 0, None, C.byref(n_mt))
    N.check(n_seg)
    cap = max(int(n_seg), 1)
    seg_query = np.zeros(cap, np.int32)
    seg_lo = np.zeros(cap, np.int32)
    seg_hi = np.zeros(cap, np.int32)
    mt_seg = np.zeros(int(n_mt.value) + 1, np.int32)
    got = lib.lis_plan_queries(lens.ctypes.data, nq, cap, seg_query.ctypes.data, seg_lo.ctypes.data,
                               seg_hi.ctypes.data, len(mt_seg), mt_seg.ctypes.data, C.byref(n_mt))
    N.check(got)
    n_seg = int(got)
    seg_query, seg_lo, seg_hi = seg_query[:n_seg], seg_lo[:n_seg], seg_hi[:n_seg]
    seg_first = np.searchsorted(seg_query, np.arange(nq + 1), side="left").astype(np.int32)
    return QueryPlan(nq, int(lens.sum()), n_seg, int(n_mt.value), seg_query, seg_lo, seg_hi, mt_seg, seg_first)


def clamp_flags(p_lens: Sequence[int], batch_size: int = 128) -> np.ndarray:
    """uint8 [np]: 1 where the reference's zero padding is visible to the page, i.e. the page is
    shorter than the longest page of its ``batch_size`` block (``pad_sequence(..., padding_value=0)``
    at HF processing_colpali.py:355-357 adds similarity-0 rows that take part in the max)."""
    lens = np.asarray(p_lens, dtype=np.int64)
    out = np.zeros(len(lens), np.uint8)
    for j in range(0, len(lens), batch_size):
        blk = lens[j:j + batch_size]
        out[j:j + batch_size] = (blk < blk.max()).astype(np.uint8)
    return out


# ------------------------------------------------------------------------------------------------
# Device-side preparation
# ------------------------------------------------------------------------------------------------
@dataclass
class PackedQueries:
    plan: QueryPlan
    rows: torch.Tensor      # [n_rows, 128] 16-bit, device: the queries' token rows back to back
    tables: torch.Tensor    # int32 device: seg_lo | seg_hi | mt_seg | seg_first
    dtype: torch.dtype      # dtype of the embeddings as given (float32 -> rows/rows_lo are bf16 planes)
    rows_lo: Optional[torch.Tensor] = None   # low plane of split-fp32 queries

    def table_ptrs(self) -> Tuple[int, int, int, int]:
        p, ns, nm = self.tables.data_ptr(), self.plan.n_seg, self.plan.n_mtiles
        return p, p + 4 * ns, p + 8 * ns, p + 8 * ns + 4 * (nm + 1)


def _as_list(x: TensorOrList) -> List[torch.Tensor]:
    if isinstance(x, torch.Tensor):
        if x.dim() != 3:
            raise ValueError(f"expected a [n, tokens, {N.DIM}] tensor, got shape {tuple(x.shape)}")
        return list(torch.unbind(x, dim=0))
    return list(x)


def _check_rows(t: torch.Tensor, what: str) -> None:
    if t.dim() != 2 or t.shape[1] != N.DIM:
        raise ValueError(f"{what}: expected [tokens, {N.DIM}], got {tuple(t.shape)}")


def _common_dtype(tensors: Sequence[torch.Tensor], what: str) -> torch.dtype:
    dt = tensors[0].dtype
    for t in tensors:
        if t.dtype != dt:
            raise ValueError(f"{what} mix dtypes {dt} and {t.dtype}")
    return dt


def pack_queries(qs: TensorOrList, device: torch.device, dtype: Optional[torch.dtype] = None) -> PackedQueries:
    """Move queries to ``device`` as one [rows, 128] matrix plus their segment tables."""
    if isinstance(qs, torch.Tensor) and qs.dim() == 3:
        nq, n_tok, d = qs.shape
        if d != N.DIM:
            raise ValueError(f"queries: embedding width {d} != {N.DIM}")
        lens = (n_tok,) * nq
        flat = qs.reshape(nq * n_tok, d)
    else:
        ql = _as_list(qs)
        for t in ql:
            _check_rows(t, "query")
        if ql:
            _common_dtype(ql, "queries")
        lens = tuple(int(t.shape[0]) for t in ql)
        if ql and all(t.device.type == "cpu" for t in ql):
            flat = torch.cat(ql, dim=0)  # one host gather, then a single H2D copy
        else:
            flat = torch.cat([t.to(device, non_blocking=True) for t in ql], dim=0) if ql else None
    plan = plan_queries(lens)
    if plan.n_seg == 0:
        raise ValueError("No queries provided")
    dt = dtype or flat.dtype
    if dt not in _DTYPES and dt != torch.float32:
        raise NotImplementedError(f"query dtype {dt}: bfloat16, float16 or float32 embeddings")
    rows = flat.to(device=device, dtype=dt, non_blocking=True).contiguous()
    if rows.data_ptr() % 16:
        rows = rows.clone()
    if dt == torch.float32:
        hi, lo = split_f32(rows)
        return PackedQueries(plan, hi, _device_tables(plan, device), dt, lo)
    return PackedQueries(plan, rows, _device_tables(plan, device), dt)


def split_f32(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 [rows,128] (CUDA) -> two bf16 planes with x = hi + lo + O(2^-18 |x|) (``lis_split_f32``)."""
    lib = N.load()
    x = x.contiguous()
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    if x.numel():
        N.check(lib.lis_split_f32(x.data_ptr(), x.shape[0], hi.data_ptr(), lo.data_ptr(), _stream(x.device)))
    return hi, lo


_TABLE_CACHE: dict = {}


def _device_tables(plan: QueryPlan, device: torch.device) -> torch.Tensor:
    """Segment tables on ``device`` (cached per plan: the upload happens once)."""
    key = (id(plan), device.index)
    hit = _TABLE_CACHE.get(key)
    if hit is not None and hit[0] is plan:
        return hit[1]
    host_tab = np.concatenate([plan.seg_lo, plan.seg_hi, plan.mt_seg, plan.seg_first]).astype(np.int32)
    tables = torch.from_numpy(host_tab).to(device)
    if len(_TABLE_CACHE) > 512:
        _TABLE_CACHE.clear()
    _TABLE_CACHE[key] = (plan, tables)
    return tables


@dataclass
class PageStore:
    """Flat ragged page-token store on one device (what ``lis_maxsim_scores`` reads)."""
    tokens: torch.Tensor            # [rows, 128] 16-bit (hi plane when the corpus is split fp32)
    offsets: torch.Tensor           # int64 [np+1]
    clamp: Optional[torch.Tensor]   # uint8 [np] or None
    n_pages: int
    tokens_lo: Optional[torch.Tensor] = None   # lo plane of a split-fp32 corpus
    dtype: Optional[torch.dtype] = None        # embedding dtype as given (defaults to tokens.dtype)

    def __post_init__(self):
        if self.dtype is None:
            self.dtype = self.tokens.dtype

    @property
    def n_rows(self) -> int:
        return int(self.tokens.shape[0])


def build_page_store(ps: TensorOrList, device: torch.device, dtype: torch.dtype, batch_size: int = 128) -> PageStore:
    if isinstance(ps, torch.Tensor) and ps.dim() == 3:
        n, s, d = ps.shape
        if d != N.DIM:
            raise ValueError(f"passages: embedding width {d} != {N.DIM}")
        tokens = ps.to(device=device, dtype=dtype, non_blocking=True).contiguous().reshape(n * s, d)
        offsets = torch.arange(0, (n + 1) * s, s, dtype=torch.int64, device=device) if s > 0 else \
            torch.zeros(n + 1, dtype=torch.int64, device=device)
        if dtype == torch.float32:
            hi, lo = split_f32(tokens)
            return PageStore(hi, offsets, None, n, lo, dtype)
        return PageStore(tokens, offsets, None, n)
    pl = _as_list(ps)
    for t in pl:
        _check_rows(t, "passage")
    _common_dtype(pl, "passages")
    lens = np.asarray([int(t.shape[0]) for t in pl], dtype=np.int64)
    host = [t for t in pl if t.device.type == "cpu"]
    if len(host) == len(pl):
        tokens = torch.cat(pl, dim=0).to(device=device, dtype=dtype, non_blocking=True)
    else:
        tokens = torch.cat([t.to(device, non_blocking=True) for t in pl], dim=0).to(dtype)
    offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(device, non_blocking=True)
    flags = clamp_flags(lens, batch_size)
    clamp = torch.from_numpy(flags).to(device, non_blocking=True) if flags.any() else None
    if dtype == torch.float32:
        hi, lo = split_f32(tokens)
        return PageStore(hi, offsets, clamp, len(pl), lo, dtype)
    return PageStore(tokens.contiguous(), offsets, clamp, len(pl))


def maxsim_scores_device(pq: PackedQueries, store: PageStore, round_mode: str = "f32",
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Run K1 (+ the segment reduction when a query was split).  Returns device fp32 [nq, np]."""
    lib = N.load()
    if round_mode not in _ROUND:
        raise ValueError(f"round_mode must be one of {sorted(_ROUND)}")
    if store.dtype != pq.dtype:
        raise ValueError(f"queries are {pq.dtype} but pages are {store.dtype}")
    device = pq.rows.device
    plan = pq.plan
    npg = store.n_pages
    seg_lo, seg_hi, mt_seg, seg_first = pq.table_ptrs()
    direct = plan.direct
    rm = _ROUND[round_mode]
    if out is None:
        out = torch.empty((plan.nq, npg), dtype=torch.float32, device=device)
    seg_out = out if direct else torch.empty((plan.n_seg, npg), dtype=torch.float32, device=device)
    st = _stream(device)
    if pq.dtype == torch.float32:
        # split-fp32 planes: torch computes fp32 inputs in fp32, so there is no rounding to emulate
        N.check(lib.lis_maxsim_scores_f32x2(pq.rows.data_ptr(), pq.rows_lo.data_ptr(), pq.rows.shape[0], seg_lo, seg_hi,
                                            mt_seg, plan.n_seg, plan.n_mtiles, store.tokens.data_ptr(),
                                            _ptr(store.tokens_lo), store.n_rows, store.offsets.data_ptr(),
                                            _ptr(store.clamp), npg, seg_out.data_ptr(), seg_out.stride(0), st))
        if not direct:
            N.check(lib.lis_reduce_segments(seg_out.data_ptr(), seg_out.stride(0), seg_first, plan.nq, npg,
                                            N.ROUND_F32, N.LIS_BF16, out.data_ptr(), out.stride(0), st))
        return out
    N.check(lib.lis_maxsim_scores(pq.rows.data_ptr(), pq.rows.shape[0], seg_lo, seg_hi, mt_seg, plan.n_seg,
                                  plan.n_mtiles, store.tokens.data_ptr(), store.n_rows, store.offsets.data_ptr(),
                                  _ptr(store.clamp), npg, _DTYPES[pq.dtype], rm if direct else rm | N.ROUND_DEFER_SUM,
                                  seg_out.data_ptr(), seg_out.stride(0), st))
    if not direct:
        N.check(lib.lis_reduce_segments(seg_out.data_ptr(), seg_out.stride(0), seg_first, plan.nq, npg,
                                        rm, _DTYPES[pq.dtype], out.data_ptr(), out.stride(0), st))
    return out


def _host_corpus(ps: TensorOrList) -> bool:
    if isinstance(ps, torch.Tensor):
        return ps.device.type == "cpu" and ps.dim() == 3
    return len(ps) > 0 and all(isinstance(t, torch.Tensor) and t.device.type == "cpu" for t in ps)


def stream_scores_host_corpus(pq: PackedQueries, ps: TensorOrList, batch_size: int, round_mode: str,
                              chunk_rows: int = 0, host_threads: int = 0) -> torch.Tensor:
    """K1 over a corpus that stays in HOST memory (``lis_stream_scores``): chunks of whole pages go through a pinned
    double buffer on a copy stream while the previous chunk is being scored -- no ``torch.cat`` / ``torch.stack`` of the
    corpus (05_experiment02.py:213 does one per call) and no need for the corpus to fit HBM.  16-bit embeddings;
    ``ps`` is a CPU ``[n, S, 128]`` tensor or a list of per-page CPU tensors.  Returns device fp32 ``[nq, np]``."""
    lib = N.load()
    device = pq.rows.device
    plan = pq.plan
    keep = None
    if isinstance(ps, torch.Tensor):
        n, s_len, d = ps.shape
        if d != N.DIM:
            raise ValueError(f"passages: embedding width {d} != {N.DIM}")
        flat = ps.detach().contiguous()
        lens = np.full(n, s_len, dtype=np.int64)
        tok_ptr, ptrs, keep = flat.data_ptr(), None, flat
    else:
        pl = list(ps)
        for t in pl:
            _check_rows(t, "passage")
        _common_dtype(pl, "passages")
        pl = [t.detach() if t.is_contiguous() else t.detach().contiguous() for t in pl]
        lens = np.asarray([int(t.shape[0]) for t in pl], dtype=np.int64)
        ptrs = np.asarray([t.data_ptr() for t in pl], dtype=np.uint64)
        tok_ptr, keep = None, pl
        n = len(pl)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n_rows = int(offsets[-1])
    if n_rows == 0:
        raise ValueError("No passages provided")
    flags = None if isinstance(ps, torch.Tensor) else clamp_flags(lens, batch_size)
    if flags is not None and not flags.any():
        flags = None
    seg_lo, seg_hi, mt_seg, seg_first = pq.table_ptrs()
    direct = plan.direct
    rm = _ROUND[round_mode]
    out = torch.empty((plan.nq, n), dtype=torch.float32, device=device)
    seg_out = out if direct else torch.empty((plan.n_seg, n), dtype=torch.float32, device=device)
    st = _stream(device)
    N.check(lib.lis_stream_scores(pq.rows.data_ptr(), pq.rows.shape[0], seg_lo, seg_hi, mt_seg, plan.n_seg, plan.n_mtiles,
                                  tok_ptr, None if ptrs is None else ptrs.ctypes.data, n_rows, offsets.ctypes.data,
                                  None if flags is None else flags.ctypes.data, n, _DTYPES[pq.dtype],
                                  rm if direct else rm | N.ROUND_DEFER_SUM, seg_out.data_ptr(), seg_out.stride(0),
                                  int(chunk_rows), int(host_threads), st))
    del keep
    if not direct:
        N.check(lib.lis_reduce_segments(seg_out.data_ptr(), seg_out.stride(0), seg_first, plan.nq, n, rm,
                                        _DTYPES[pq.dtype], out.data_ptr(), out.stride(0), st))
    return out


def calibrate_pass_costs(device: Union[str, torch.device, None] = None, pages: int = 12_000, page_tokens: int = 1030,
                         iters: int = 6) -> dict:
    """Measure, on THIS device, what one pass of every K1 form costs (one CTA per SM with 1..3 resident query tiles, CTA
    pairs with 2..10) over a synthetic ColPali-shaped store, and install the table in the pass planner
    (``lis_set_pass_costs``).  The built-in table comes from one power-capped B200; boxes differ by +-10 %.  Takes about a
    second; returns ``{"single": [...], "pair": [...]}`` in milliseconds."""
    from .index import LateInteractionIndex

    lib = N.load()
    dev = resolve_device(device)
    idx = LateInteractionIndex(pages * page_tokens, pages, device=dev)
    single, pair = np.zeros(4, np.float32), np.zeros(11, np.float32)
    try:
        idx.fill_synthetic(pages, page_tokens, seed=11)
        store = idx._as_store()
        g = torch.Generator().manual_seed(1)
        for form, counts, table in (("single", (1, 2, 3), single), ("pair", tuple(range(2, 11)), pair)):
            for n in counts:
                q = torch.nn.functional.normalize(torch.randn(n * 4, 32, N.DIM, generator=g), dim=-1).to(torch.bfloat16).to(dev)
                pq = pack_queries(q, dev)
                out = torch.empty((n * 4, pages), dtype=torch.float32, device=dev)
                N.check(lib.lis_set_tuning(0, n, 0, 0, 1 if form == "single" else 3))
                for _ in range(3):
                    maxsim_scores_device(pq, store, "f32", out=out)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    maxsim_scores_device(pq, store, "f32", out=out)
                e1.record()
                torch.cuda.synchronize(dev)
                table[n] = e0.elapsed_time(e1) / iters
    finally:
        lib.lis_set_tuning(0, 0, 0, 0, 0)
        idx.close()
    N.check(lib.lis_set_pass_costs(single.ctypes.data, pair.ctypes.data))
    return {"single": single.tolist(), "pair": pair.tolist()}


_COPY_STREAMS: dict = {}


def scores_to_host_overlapped(pq: PackedQueries, store: PageStore, round_mode: str, out: torch.Tensor,
                              n_blocks: int = 4) -> torch.Tensor:
    """K1 over `n_blocks` ranges of pages; the score columns of a finished range travel to the host result (`out`, CPU
    float32 ``[nq, np]``, ideally pinned) on a copy stream while the next range is being scored -- the device-to-host
    leg that colpali-engine pays per 128-page block with ``.cpu()`` (HF processing_colpali.py:362) hides behind the kernel."""
    lib = N.load()
    device = pq.rows.device
    npg = store.n_pages
    scores = torch.empty((pq.plan.nq, npg), dtype=torch.float32, device=device)
    copy = _COPY_STREAMS.get(device.index)
    if copy is None:
        copy = _COPY_STREAMS[device.index] = torch.cuda.Stream(device)
    main = torch.cuda.current_stream(device)
    bounds = [npg * i // n_blocks for i in range(n_blocks + 1)]
    for a, b in zip(bounds[:-1], bounds[1:]):
        if b <= a:
            continue
        sub = PageStore(store.tokens, store.offsets[a:b + 1], None if store.clamp is None else store.clamp[a:b], b - a,
                        store.tokens_lo, store.dtype)
        maxsim_scores_device(pq, sub, round_mode, out=scores[:, a:b])
        ev = torch.cuda.Event()
        ev.record(main)
        copy.wait_event(ev)
        N.check(lib.lis_memcpy2d_async(out.data_ptr() + 4 * a, out.stride(0) * 4, scores.data_ptr() + 4 * a, scores.stride(0) * 4,
                                       4 * (b - a), pq.plan.nq, copy.cuda_stream))
    copy.synchronize()
    scores.record_stream(copy)
    return out


def score_multi_vector(qs: TensorOrList, ps: TensorOrList, batch_size: int = 128,
                       device: Union[str, torch.device, None] = None, *, round_mode: str = "reference",
                       return_device: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``score[b, c] = sum_n max_s <qs[b][n], ps[c][s]>`` -- same signature, zero-padding semantics,
    error behaviour and CPU-float32 ``[len(qs), len(ps)]`` result as colpali-engine's
    ``score_multi_vector`` (05_experiment02.py:214).

    ``round_mode="reference"`` (default) reproduces what torch does to 16-bit inputs (per-token max
    and final sum rounded to the input dtype); ``"f32"`` keeps fp32 throughout (the more accurate
    number; within 1e-4 of the fp32-widened reference).  float32 embeddings (ColFlor's default,
    05_experiment02.py:343-347) are scored as two bf16 planes per operand with fp32 accumulation
    (error ~1e-6 on unit-norm rows, inside the 1e-4 bar); ``round_mode`` does not apply to them.  ``batch_size`` only matters through the
    padding it implies in the reference: a page shorter than the longest page of its block has its
    per-token max clamped at 0.  Inputs may live on any device; the corpus is not copied when it is
    already a contiguous CUDA tensor.  ``out``: optional CPU float32 ``[len(qs), len(ps)]`` tensor to
    receive the scores (pin it to make the device-to-host copy DMA-direct).
    """
    if len(qs) == 0:
        raise ValueError("No queries provided")
    if len(ps) == 0:
        raise ValueError("No passages provided")
    q_dt = qs.dtype if isinstance(qs, torch.Tensor) else qs[0].dtype
    p_dt = ps.dtype if isinstance(ps, torch.Tensor) else ps[0].dtype
    if q_dt != p_dt:   # torch.einsum in the reference refuses mixed dtypes too (HF port: processing_colpali.py:342)
        raise ValueError(f"Queries and passages must have the same dtype (queries are {q_dt}, passages {p_dt})")
    dev = resolve_device(device)
    N.check(N.load().lis_device_supported(dev.index))
    with torch.cuda.device(dev):
        pq = pack_queries(qs, dev)
        if round_mode not in _ROUND:
            raise ValueError(f"round_mode must be one of {sorted(_ROUND)}")
        if pq.dtype in _DTYPES and _host_corpus(ps):
            # the reference's literal call: the corpus is a CPU tensor / list (05_experiment02.py:213-214)
            scores = stream_scores_host_corpus(pq, ps, batch_size, round_mode)
        else:
            store = build_page_store(ps, dev, pq.dtype, batch_size)
            if store.n_rows == 0:
                raise ValueError("No passages provided")
            if not return_device and store.n_pages >= 8192:
                # large result: send finished column blocks to the host while the next pages are being scored
                if out is None:
                    out = torch.empty((pq.plan.nq, store.n_pages), dtype=torch.float32)
                elif out.device.type != "cpu" or out.dtype != torch.float32 or tuple(out.shape) != (pq.plan.nq, store.n_pages) \
                        or out.stride(1) != 1:
                    raise ValueError(f"out must be a CPU float32 tensor of shape {(pq.plan.nq, store.n_pages)}")
                return scores_to_host_overlapped(pq, store, round_mode, out)
            scores = maxsim_scores_device(pq, store, round_mode)
        if return_device:
            return scores
        if out is not None:
            if out.device.type != "cpu" or out.dtype != torch.float32 or tuple(out.shape) != tuple(scores.shape):
                raise ValueError(f"out must be a CPU float32 tensor of shape {tuple(scores.shape)}")
            out.copy_(scores, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return out
        return scores.cpu()
