"""Generate the committed golden vectors for the scoring path.

The reference's own scoring body (colpali-engine 0.3.13 ``score_multi_vector``) is not installable
offline; the arithmetically identical port that IS installed in this image is
``transformers.models.colpali.processing_colpali.ColPaliProcessor.score_retrieval`` (transformers
5.5.0, processing_colpali.py:302-364).  This script runs THAT function (not the oracle) on seeded
inputs and freezes inputs + outputs, so tests on a machine without it still pin the oracle.
The projection-head vectors come from the literal statements of HF modeling_colpali.py:148-155
executed with ``torch.nn.functional.linear``.

    python tests/golden/make_golden.py      # rewrites tests/golden/maxsim_golden.pt
"""
from pathlib import Path

import torch
from transformers.models.colpali.processing_colpali import ColPaliProcessor

OUT = Path(__file__).resolve().parent / "maxsim_golden.pt"


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def hf_scores(qs, ps, batch_size=128):
    # score_retrieval never touches `self`
    return ColPaliProcessor.score_retrieval(None, qs, ps, batch_size=batch_size, output_dtype=torch.float32)


def main():
    g = torch.Generator().manual_seed(20261018)
    cases = {}
    # (a) padded tensors, ColPali-like but tiny
    q = unit(torch.randn(3, 7, 128, generator=g)).to(torch.bfloat16)
    p = unit(torch.randn(9, 21, 128, generator=g)).to(torch.bfloat16)
    cases["padded"] = dict(qs=q, ps=p, batch_size=128)
    # (b) ragged lists crossing 128-page block boundaries (zero-padding semantics inside each block)
    q_lens = [1, 4, 9, 16, 2]
    p_lens = [int(x) for x in torch.randint(1, 13, (260,), generator=g)]
    cases["ragged"] = dict(qs=[unit(torch.randn(n, 128, generator=g)).to(torch.bfloat16) for n in q_lens],
                           ps=[unit(torch.randn(n, 128, generator=g)).to(torch.bfloat16) for n in p_lens],
                           batch_size=128)
    # (c) same data, small batch size: different padding blocks, different clamping
    cases["ragged_bs8"] = dict(qs=cases["ragged"]["qs"], ps=cases["ragged"]["ps"][:40], batch_size=8)
    # (d) anti-aligned short page: the zero pad rows win the max
    qd = unit(torch.randn(1, 6, 128, generator=g)).to(torch.bfloat16)
    cases["negative"] = dict(qs=qd, ps=[(-qd[0, :2]).clone(), unit(torch.randn(11, 128, generator=g)).to(torch.bfloat16)],
                             batch_size=128)
    out = {}
    for name, c in cases.items():
        qs, ps, bs = c["qs"], c["ps"], c["batch_size"]
        widen = (lambda x: x.float()) if isinstance(qs, torch.Tensor) else (lambda x: [t.float() for t in x])
        widen_p = (lambda x: x.float()) if isinstance(ps, torch.Tensor) else (lambda x: [t.float() for t in x])
        out[name] = dict(qs=qs, ps=ps, batch_size=bs,
                         scores_bf16=hf_scores(qs, ps, bs),                 # the reference's 16-bit path
                         scores_fp32=hf_scores(widen(qs), widen_p(ps), bs))  # same inputs widened to fp32
    # projection head (HF modeling_colpali.py:148-155)
    h = torch.randn(2, 10, 192, generator=g)
    w = torch.randn(128, 192, generator=g) / 192 ** 0.5
    b = 0.1 * torch.randn(128, generator=g)
    mask = (torch.rand(2, 10, generator=g) > 0.3).long()
    emb = torch.nn.functional.linear(h, w, b)
    emb = emb / emb.norm(dim=-1, keepdim=True)
    emb = emb * mask.unsqueeze(-1)
    out["head"] = dict(hidden=h, weight=w, bias=b, mask=mask, embeddings=emb)
    torch.save(out, OUT)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
