"""TEST INFRASTRUCTURE ONLY -- CPU oracle for late-interaction (MaxSim) scoring.

What it restates
----------------
The reference calls ``processor.score_multi_vector(qs, ps)`` at
``05_experiment02.py:214`` (notebook twin ``05_experiment02.ipynb:228``).  The
body of that method is NOT in the reference tree: it lives in the third-party
dependency ``colpali-engine==0.3.13`` (``poetry.lock:795-803``,
``pyproject.toml:28``), class ``BaseVisualRetrieverProcessor``.  That package is
not installable here (no network).  Its published algorithm is restated below
and is pinned against the arithmetically identical port that IS installed in
this image: ``transformers 5.5.0``
``ColPaliProcessor.score_retrieval`` (``processing_colpali.py:302-364``, core
expression at ``:360``).  ``tests/test_oracle.py`` checks bit-equality against
that port, and ``tests/golden/make_golden.py`` froze its outputs as fixtures.

Parity status: the reference repository has no tests, golden vectors or
fixtures of its own for this path ("parity unpinned" by the reference itself,
SURVEY.md section 4); the pin used here is the installed HF port plus analytic
known-answer cases.

Algorithm (colpali-engine 0.3.13 ``score_multi_vector``):
    for i in range(0, len(qs), batch_size):                     # 128 queries
        qb = pad_sequence(qs[i:i+bs], batch_first=True, padding_value=0).to(device)
        for j in range(0, len(ps), batch_size):                 # 128 pages
            pb = pad_sequence(ps[j:j+bs], batch_first=True, padding_value=0).to(device)
            block = einsum("bnd,csd->bcns", qb, pb).max(dim=3)[0].sum(dim=2)
        row = cat(blocks, dim=1).cpu()
    scores = cat(rows, dim=0).to(float32)                       # CPU fp32 [nq, np]
Empty ``qs`` / ``ps`` raise ``ValueError``.

Zero padding is part of the semantics: a page shorter than the longest page
of its 128-page block sees similarity exactly 0 from the padded rows, so its
per-token max is clamped at >= 0 (hard part 4).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple, Union

import torch

TensorOrList = Union[torch.Tensor, Sequence[torch.Tensor]]


def _as_list(x: TensorOrList) -> List[torch.Tensor]:
    if isinstance(x, torch.Tensor):
        return list(torch.unbind(x, dim=0))
    return list(x)


def score_multi_vector(qs: TensorOrList, ps: TensorOrList, batch_size: int = 128,
                       device: Union[str, torch.device] = "cpu") -> torch.Tensor:
    """Restatement of colpali-engine 0.3.13 ``score_multi_vector`` (call site
    05_experiment02.py:214; arithmetic identical to HF processing_colpali.py:350-364).
    Computes in the input dtype, returns CPU float32 ``[len(qs), len(ps)]``."""
    if len(qs) == 0:
        raise ValueError("No queries provided")
    if len(ps) == 0:
        raise ValueError("No passages provided")
    qs = _as_list(qs)
    ps = _as_list(ps)
    rows: List[torch.Tensor] = []
    for i in range(0, len(qs), batch_size):
        qb = torch.nn.utils.rnn.pad_sequence(qs[i:i + batch_size], batch_first=True,
                                             padding_value=0).to(device)
        blocks: List[torch.Tensor] = []
        for j in range(0, len(ps), batch_size):
            pb = torch.nn.utils.rnn.pad_sequence(ps[j:j + batch_size], batch_first=True,
                                                 padding_value=0).to(device)
            blocks.append(torch.einsum("bnd,csd->bcns", qb, pb).max(dim=3)[0].sum(dim=2))
        rows.append(torch.cat(blocks, dim=1).cpu())
    scores = torch.cat(rows, dim=0)
    assert scores.shape[0] == len(qs)
    return scores.to(torch.float32)


def score_multi_vector_widened(qs: TensorOrList, ps: TensorOrList, batch_size: int = 128,
                               dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Accuracy oracle: the same inputs widened to fp32 (or fp64) before the
    restated arithmetic.  The 1e-4 tolerance in BASELINE.json refers to this."""
    qs = [q.to(dtype) for q in _as_list(qs)]
    ps = [p.to(dtype) for p in _as_list(ps)]
    return score_multi_vector(qs, ps, batch_size=batch_size)


def score_multi_vector_bf16_rounding_model(qs: TensorOrList, ps: TensorOrList,
                                           batch_size: int = 128) -> torch.Tensor:
    """Model of what torch does to bf16 inputs (SURVEY.md header note 3):
    ``bf16( sum_n fp32( bf16( max_s fp32dot ) ) )``.  Used to check the kernel's
    ``round_mode=reference`` epilogue without depending on a BLAS's accumulation
    order.  Inputs must be bf16 (or are rounded to bf16 first)."""
    qs = [q.to(torch.bfloat16).to(torch.float32) for q in _as_list(qs)]
    ps = [p.to(torch.bfloat16).to(torch.float32) for p in _as_list(ps)]
    if len(qs) == 0:
        raise ValueError("No queries provided")
    if len(ps) == 0:
        raise ValueError("No passages provided")
    rows = []
    for i in range(0, len(qs), batch_size):
        qb = torch.nn.utils.rnn.pad_sequence(qs[i:i + batch_size], batch_first=True)
        blocks = []
        for j in range(0, len(ps), batch_size):
            pb = torch.nn.utils.rnn.pad_sequence(ps[j:j + batch_size], batch_first=True)
            sim = torch.einsum("bnd,csd->bcns", qb, pb)
            mx = sim.max(dim=3)[0].to(torch.bfloat16).to(torch.float32)
            blocks.append(mx.sum(dim=2).to(torch.bfloat16))
        rows.append(torch.cat(blocks, dim=1))
    return torch.cat(rows, dim=0).to(torch.float32)


def topk(scores: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-query top-k, restating ``query_scores.topk(top_k)`` (05_experiment02.py:219)
    with the tie rule torch leaves unspecified made explicit: score descending,
    then page index ascending.  Returns (values [nq,k] fp32, indices [nq,k] int64)."""
    scores = scores.to(torch.float32)
    nq, n = scores.shape
    k = min(k, n)
    # stable sort on descending score keeps ascending index order inside ties
    order = torch.sort(scores, dim=1, descending=True, stable=True).indices[:, :k]
    return torch.gather(scores, 1, order), order


def merge_topk(parts: Sequence[Tuple[torch.Tensor, torch.Tensor]], k: int
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge per-shard (values [nq,k_i], global ids [nq,k_i]) candidate lists into a
    global top-k with the same (score desc, id asc) rule.  Entries with id < 0 are
    padding and sort last."""
    vals = torch.cat([p[0].to(torch.float32) for p in parts], dim=1)
    ids = torch.cat([p[1].to(torch.int64) for p in parts], dim=1)
    out_v, out_i = [], []
    for r in range(vals.shape[0]):
        rows = [(-(float(v)), int(i)) for v, i in zip(vals[r].tolist(), ids[r].tolist()) if i >= 0]
        rows.sort()
        rows = rows[:k]
        out_v.append([-a for a, _ in rows] + [float("-inf")] * (k - len(rows)))
        out_i.append([b for _, b in rows] + [-1] * (k - len(rows)))
    return torch.tensor(out_v, dtype=torch.float32), torch.tensor(out_i, dtype=torch.int64)


def project_normalize(hidden: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None,
                      attention_mask: torch.Tensor | None) -> torch.Tensor:
    """Restatement of the retrieval head the reference runs inside ``model(**batch)``
    (functions.py:795, 839, 888; 05_experiment02.py:211).  Body follows HF
    modeling_colpali.py:148-155: Linear -> x / ||x||_2 (no epsilon) -> * mask.
    Computed in the dtype of ``weight``; returns that dtype."""
    x = torch.nn.functional.linear(hidden.to(weight.dtype), weight, bias)
    x = x / x.norm(dim=-1, keepdim=True)
    if attention_mask is not None:
        x = x * attention_mask.unsqueeze(-1)
    return x
