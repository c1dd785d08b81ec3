"""Concurrent single-query batching.

A search over a large corpus is HBM-bound: one pass costs the same whether it serves 1 query or a
full 128-row M tile of query tokens (and barely more for two tiles).  ``QueryBatcher`` coalesces the
searches that are in flight -- e.g. the per-question ``RetrievalManager.fetch`` calls of
02_experiment01.py:141-164, or the asyncio fan-out of 05_experiment02.py:297-298 -- into one pass,
multiplying queries/s at unchanged per-pass latency.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from typing import List, Tuple

import torch


class QueryBatcher:
    def __init__(self, index, max_rows: int = 256, max_wait_ms: float = 0.2, round_mode: str = "f32"):
        """``index``: anything with ``search(list_of_queries, k, round_mode) -> (scores, ids)``
        (:class:`LateInteractionIndex`, :class:`ShardedIndex`).  ``max_rows``: query-token rows per pass
        (256 = two M tiles stays on the HBM roofline).  ``max_wait_ms``: how long the first query of a
        batch may wait for company."""
        self.index = index
        self.max_rows = int(max_rows)
        self.max_wait = float(max_wait_ms) * 1e-3
        self.round_mode = round_mode
        self._q: "queue.Queue" = queue.Queue()
        self._stop = threading.Event()
        self.batches = 0
        self.served = 0
        self._worker = threading.Thread(target=self._run, name="lis-query-batcher", daemon=True)
        self._worker.start()

    def submit(self, query: torch.Tensor, k: int) -> Future:
        """``query`` [n_tok, 128]; resolves to (scores [k], ids [k]) on the CPU."""
        if query.dim() != 2:
            raise ValueError("submit() takes one query of shape [n_tok, 128]")
        fut: Future = Future()
        self._q.put((query, int(k), fut))
        return fut

    def search(self, query: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.submit(query, k).result()

    def close(self) -> None:
        self._stop.set()
        self._q.put(None)
        self._worker.join(timeout=5)

    def _run(self) -> None:
        while not self._stop.is_set():
            item = self._q.get()
            if item is None:
                break
            batch: List = [item]
            rows = item[0].shape[0]
            deadline = time.perf_counter() + self.max_wait
            while rows < self.max_rows:
                left = deadline - time.perf_counter()
                try:
                    nxt = self._q.get(timeout=left) if left > 0 else self._q.get_nowait()
                except queue.Empty:
                    break
                if nxt is None:
                    self._stop.set()
                    break
                if rows + nxt[0].shape[0] > self.max_rows and len(batch) > 0:
                    self._q.put(nxt)      # keep it for the next pass
                    break
                batch.append(nxt)
                rows += nxt[0].shape[0]
            try:
                kmax = max(b[1] for b in batch)
                scores, ids = self.index.search([b[0] for b in batch], kmax, self.round_mode)
                for i, (_, k, fut) in enumerate(batch):
                    fut.set_result((scores[i, :k].clone(), ids[i, :k].clone()))
            except BaseException as exc:  # propagate to every waiter
                for _, _, fut in batch:
                    if not fut.done():
                        fut.set_exception(exc)
            self.batches += 1
            self.served += len(batch)
