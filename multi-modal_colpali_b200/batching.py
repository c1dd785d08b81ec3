"""Concurrent single-query batching.

A search over a large corpus is HBM-bound: one pass costs the same whether it serves 1 query or a
full 128-row M tile of query tokens (and barely more for two tiles).  ``QueryBatcher`` coalesces the
searches that are in flight -- e.g. the per-question ``RetrievalManager.fetch`` calls of
02_experiment01.py:141-164, or the asyncio fan-out of 05_experiment02.py:297-298 -- into one pass,
multiplying queries/s at unchanged per-pass latency.

Results are those of ``index.search([query], k)`` bit for bit: K1 scores every (query, page) pair independently
of what else shares the pass -- the per-token maxima are exact, and the sum over a query's tokens follows one
canonical order per segment (``reduce_tile_segments``) -- provided the query is cut into the same segments as when it
is searched alone.  Segments are cut at multiples of 64 packed rows, so the batcher places every query where it
crosses none (or, for a query longer than 64 tokens, where it starts on one), padding with zero rows when needed.
The top-``k`` prefix of a top-``kmax`` list under the total order (score desc, id asc) is the top-``k`` list.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from typing import List, Optional, Tuple

import torch

from . import _native as N


class QueryBatcher:
    def __init__(self, index, max_rows: int = 256, max_wait_ms: float = 0.2, round_mode: str = "f32"):
        """``index``: anything with ``search(list_of_queries, k, round_mode) -> (scores, ids)``
        (:class:`LateInteractionIndex`, :class:`ShardedIndex`).  ``max_rows``: query-token rows per pass
        (256 = two M tiles stays on the HBM roofline).  ``max_wait_ms``: how long the first query of a
        batch may wait for company.  One worker thread owns the index: it is the only caller of
        ``index.search``, which is what the index's one-search-in-flight rule asks for."""
        self.index = index
        self.max_rows = int(max_rows)
        self.max_wait = float(max_wait_ms) * 1e-3
        self.round_mode = round_mode
        self._dtype = getattr(index, "dtype", None)
        self._q: "queue.Queue" = queue.Queue()
        self._closed = False
        self._lock = threading.Lock()
        self.batches = 0
        self.served = 0
        self._worker = threading.Thread(target=self._run, name="lis-query-batcher", daemon=True)
        self._worker.start()

    def submit(self, query: torch.Tensor, k: int) -> Future:
        """``query`` [n_tok, 128]; resolves to (scores [k], ids [k]) on the CPU.  Malformed requests fail here,
        individually, and never reach a coalesced batch."""
        if not isinstance(query, torch.Tensor) or query.dim() != 2 or query.shape[1] != N.DIM:
            raise ValueError(f"submit() takes one query of shape [n_tok, {N.DIM}]")
        if query.shape[0] == 0:
            raise ValueError("No queries provided")
        if query.shape[0] > self.max_rows:
            raise ValueError(f"query has {query.shape[0]} token rows; this batcher coalesces up to {self.max_rows}")
        k = int(k)
        if not 1 <= k <= N.MAX_K:
            raise ValueError(f"k={k} out of range 1..{N.MAX_K}")
        if not query.is_floating_point():
            raise ValueError("query must be a floating-point tensor")
        if self._dtype is not None and query.dtype != self._dtype:
            query = query.to(self._dtype)      # what index.search would do to a lone query
        fut: Future = Future()
        with self._lock:
            if self._closed:
                raise RuntimeError("QueryBatcher is closed")
            self._q.put((query, k, fut))
        return fut

    def search(self, query: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.submit(query, k).result()

    def close(self) -> None:
        """Stop accepting work; everything submitted before the call is still served."""
        with self._lock:
            if self._closed:
                return
            self._closed = True
            self._q.put(None)
        self._worker.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    _CUT = 64      # lis_plan_queries cuts segments at multiples of 64 packed rows

    @classmethod
    def _placed_rows(cls, rows: int, n_tok: int) -> int:
        """Packed rows in use after appending a query of ``n_tok`` tokens behind ``rows`` rows, padding included."""
        room = cls._CUT - rows % cls._CUT
        if rows % cls._CUT and (n_tok > room):
            rows += room           # start on the next cut: the segments then equal those of the query searched alone
        return rows + n_tok

    def _serve(self, batch: List) -> None:
        try:
            kmax = max(b[1] for b in batch)
            qs, where, rows = [], [], 0
            for q, _, _ in batch:
                room = self._CUT - rows % self._CUT
                if rows % self._CUT and q.shape[0] > room:
                    qs.append(torch.zeros((room, N.DIM), dtype=q.dtype, device=q.device))   # filler query, result ignored
                    rows += room
                where.append(len(qs))
                qs.append(q)
                rows += q.shape[0]
            scores, ids = self.index.search(qs, kmax, self.round_mode)
            for (_, k, fut), i in zip(batch, where):
                fut.set_result((scores[i, :k].clone(), ids[i, :k].clone()))
        except BaseException as exc:
            if len(batch) == 1:
                batch[0][2].set_exception(exc)
            else:                      # do not let one request fail its neighbours: retry them one at a time
                for item in batch:
                    self._serve([item])
                return
        self.batches += 1
        self.served += len(batch)

    def _run(self) -> None:
        pending: Optional[tuple] = None      # a request that did not fit the previous pass: head of the next one
        draining = False
        while True:
            item = pending if pending is not None else self._q.get()
            pending = None
            if item is None:
                break                        # the sentinel is the last thing ever queued (submit refuses after close)
            batch: List = [item]
            rows = item[0].shape[0]       # packed rows incl. the padding that keeps queries off the 64-row cuts
            deadline = time.perf_counter() + self.max_wait
            while rows < self.max_rows:
                left = deadline - time.perf_counter()
                try:
                    nxt = self._q.get(timeout=left) if (left > 0 and not draining) else self._q.get_nowait()
                except queue.Empty:
                    break
                if nxt is None:
                    draining = True
                    pending = None
                    self._q.put(None)        # keep the sentinel at the tail; nothing can follow it
                    break
                if self._placed_rows(rows, nxt[0].shape[0]) > self.max_rows:
                    pending = nxt            # keeps its place in the order of arrival
                    break
                batch.append(nxt)
                rows = self._placed_rows(rows, nxt[0].shape[0])
            self._serve(batch)
