"""Pin the oracle: (1) bit-equality with the installed HF port of the same arithmetic, (2) the
committed golden vectors that port produced, (3) analytic known answers, (4) the bf16 rounding model.
The reference repository itself has no tests for this path ("parity unpinned" by the reference)."""
from pathlib import Path

import pytest
import torch

GOLDEN = Path(__file__).resolve().parent / "golden" / "maxsim_golden.pt"


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=True)


def _widen(x):
    return x.float() if isinstance(x, torch.Tensor) else [t.float() for t in x]


@pytest.mark.parametrize("case", ["padded", "ragged", "ragged_bs8", "negative"])
def test_oracle_matches_golden(oracle, golden, case):
    c = golden[case]
    got16 = oracle.score_multi_vector(c["qs"], c["ps"], batch_size=c["batch_size"])
    got32 = oracle.score_multi_vector_widened(c["qs"], c["ps"], batch_size=c["batch_size"])
    assert got16.dtype == torch.float32 and got32.dtype == torch.float32
    assert torch.equal(got16, c["scores_bf16"])
    assert torch.equal(got32, c["scores_fp32"])


def test_oracle_matches_installed_hf_port_live(oracle):
    mod = pytest.importorskip("transformers.models.colpali.processing_colpali")
    g = torch.Generator().manual_seed(123)
    for dtype in (torch.float32, torch.bfloat16):
        qs = [unit(torch.randn(n, 128, generator=g)).to(dtype) for n in (3, 17, 1, 40)]
        ps = [unit(torch.randn(int(n), 128, generator=g)).to(dtype) for n in torch.randint(1, 60, (150,), generator=g)]
        for bs in (128, 32):
            want = mod.ColPaliProcessor.score_retrieval(None, qs, ps, batch_size=bs, output_dtype=torch.float32)
            assert torch.equal(oracle.score_multi_vector(qs, ps, batch_size=bs), want)


def test_bf16_rounding_model_reproduces_torch_bf16(oracle, golden):
    for case in ("padded", "ragged", "negative"):
        c = golden[case]
        model = oracle.score_multi_vector_bf16_rounding_model(c["qs"], c["ps"], batch_size=c["batch_size"])
        assert torch.equal(model, c["scores_bf16"]), case


def test_known_answers(oracle):
    eye = torch.eye(128)
    q = eye[:10].unsqueeze(0)
    pages = [eye[:64], eye[5:69], eye[64:]]
    assert oracle.score_multi_vector(q, pages).tolist() == [[10.0, 5.0, 0.0]]
    q0 = torch.cat([q, torch.zeros(1, 1, 128)], dim=1)
    assert torch.equal(oracle.score_multi_vector(q0, pages), oracle.score_multi_vector(q, pages))
    g = torch.Generator().manual_seed(1)
    p = unit(torch.randn(5, 30, 128, generator=g))
    qq = unit(torch.randn(2, 9, 128, generator=g))
    perm = p[:, torch.randperm(30, generator=g)]
    assert torch.allclose(oracle.score_multi_vector(qq, perm), oracle.score_multi_vector(qq, p), atol=1e-6)
    # a page that contains the query's own tokens scores exactly the number of tokens
    page = torch.cat([qq[0], p[0]], dim=0)
    assert oracle.score_multi_vector(qq[:1], [page]).item() == pytest.approx(9.0, abs=1e-5)


def test_zero_padding_clamps_short_pages(oracle, golden):
    c = golden["negative"]
    s = c["scores_fp32"]
    alone = oracle.score_multi_vector_widened(c["qs"], [c["ps"][0]])
    assert alone.item() < s[0, 0].item()      # padded next to a longer page: the pad rows' 0 wins


def test_empty_inputs(oracle):
    x = torch.zeros(1, 2, 128)
    with pytest.raises(ValueError, match="No queries provided"):
        oracle.score_multi_vector([], x)
    with pytest.raises(ValueError, match="No passages provided"):
        oracle.score_multi_vector(x, [])


def test_topk_and_merge_rules(oracle):
    s = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0]])
    v, i = oracle.topk(s, 3)
    assert i.tolist() == [[1, 2, 4]] and v.tolist() == [[3.0, 3.0, 3.0]]
    a = (torch.tensor([[5.0, 1.0]]), torch.tensor([[7, 3]]))
    b = (torch.tensor([[5.0, float("-inf")]]), torch.tensor([[2, -1]]))
    mv, mi = oracle.merge_topk([a, b], 3)
    assert mi.tolist() == [[2, 7, 3]] and mv.tolist() == [[5.0, 5.0, 1.0]]
    mv, mi = oracle.merge_topk([a, b], 4)
    assert mi.tolist() == [[2, 7, 3, -1]]


def test_head_matches_golden(oracle, golden):
    h = golden["head"]
    got = oracle.project_normalize(h["hidden"], h["weight"], h["bias"], h["mask"])
    assert torch.equal(got, h["embeddings"])
