"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star):
* ``round_mode="f32"``: 1e-4 absolute against the oracle on the same 16-bit inputs widened to fp32.
* ``round_mode="reference"`` against the oracle run in bf16 like the reference does: 1e-2 absolute
  wherever bf16 can express it, i.e. for |score| < 2 (bf16 spacing <= 2^-7); above that two bf16
  results that differ at all differ by a whole bf16 step (2^-5 = 0.031 on [4, 8)), so the bound is
  ONE bf16 step of the score -- which is also the gap between torch's own CPU and GPU bf16 paths --
  plus a floor on the share of bit-identical entries.  A 1-step flip happens when an fp32 dot lands
  within accumulation-order noise of a bf16 rounding boundary of the per-token max.
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-4
TOL_BF16 = 1e-2


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def rand_unit(gen, *shape, dtype=torch.bfloat16):
    return unit(torch.randn(*shape, generator=gen)).to(dtype)


def ragged(gen, lens, dtype=torch.bfloat16):
    return [rand_unit(gen, int(n), 128, dtype=dtype) if n > 0 else torch.zeros(0, 128, dtype=dtype) for n in lens]


def bf16_step(x):
    """Spacing of bf16 at |x| (8 significant bits)."""
    return torch.exp2(torch.floor(torch.log2(x.abs().clamp_min(1e-30))) - 7)


def assert_topk_equiv(got_ids, got_scores, full_scores, tol):
    """north_star: top-k ids identical to the oracle's except for ties inside the tolerance."""
    k = len(got_ids)
    best = torch.sort(full_scores, descending=True).values[:k]
    assert len(set(got_ids)) == k
    for j, (i, s) in enumerate(zip(got_ids, got_scores)):
        assert abs(full_scores[i].item() - s) <= tol, (j, i, s, full_scores[i].item())
        assert abs(best[j].item() - s) <= tol, (j, i, s, best[j].item())


def check_both_modes(lis, oracle, qs, ps, batch_size=128):
    want32 = oracle.score_multi_vector_widened(qs, ps, batch_size=batch_size)
    got32 = lis.score_multi_vector(qs, ps, batch_size=batch_size, round_mode="f32")
    assert got32.dtype == torch.float32 and got32.device.type == "cpu"
    assert got32.shape == want32.shape
    err32 = (got32 - want32).abs().max().item()
    assert err32 <= TOL_F32, f"f32 mode: max abs err {err32}"
    # the reference's own 16-bit path (torch CPU) and its rounding model
    want16 = oracle.score_multi_vector(qs, ps, batch_size=batch_size)
    got16 = lis.score_multi_vector(qs, ps, batch_size=batch_size)  # default round_mode="reference"
    diff16 = (got16 - want16).abs()
    err16 = diff16.max().item()
    step = torch.maximum(bf16_step(want16), torch.tensor(TOL_BF16))
    assert (diff16 <= step).all(), f"reference mode: max abs err {err16} exceeds one bf16 step"
    assert (diff16 == 0).float().mean().item() >= 0.95, "reference mode: too few bit-identical scores"
    return err32, err16


# ---------------------------------------------------------------------------------------------
def test_config1_colpali_shapes(lis, oracle):
    """BASELINE configs[0]: 1 query x 16 tokens vs 1000 pages x 1030 tokens (seeds from SURVEY 8d)."""
    q = rand_unit(torch.Generator().manual_seed(1001), 1, 16, 128)
    p = rand_unit(torch.Generator().manual_seed(2001), 1000, 1030, 128)
    err32, err16 = check_both_modes(lis, oracle, q, p)
    # reference-rounding mode should reproduce torch's bf16 path essentially bit-for-bit
    want = oracle.score_multi_vector_bf16_rounding_model(q, p)
    got = lis.score_multi_vector(q, p)
    ulp = 2.0 ** (math.floor(math.log2(want.abs().max().item())) - 7)
    assert (got - want).abs().max().item() <= ulp
    assert ((got - want) == 0).float().mean().item() > 0.98


def test_many_queries_multi_mtile(lis, oracle):
    """32 queries x 20 tokens (BASELINE configs[1] query shape: 640 rows = 5 M tiles, cut queries)."""
    q = rand_unit(torch.Generator().manual_seed(1002), 32, 20, 128)
    p = rand_unit(torch.Generator().manual_seed(2002), 600, 257, 128)
    check_both_modes(lis, oracle, q, p)


def test_reference_mode_vs_torch_gpu_route(lis, oracle):
    """The reference's loop run on the SAME GPU through torch (cuBLAS einsum + max + sum in bf16, what colpali-engine does
    with device="cuda") against the fused kernel in round_mode="reference": within one bf16 step, >= 95 % bit-identical
    (on the round-1 boxes all scores were bit-identical)."""
    g = torch.Generator().manual_seed(77)
    q = rand_unit(g, 32, 20, 128)
    p = ragged(g, [1030] * 300 + [int(x) for x in torch.randint(200, 900, (212,), generator=g)])
    want = oracle.score_multi_vector(q, p, device="cuda")
    got = lis.score_multi_vector(q, p)
    diff = (got - want).abs()
    step = torch.maximum(bf16_step(want), torch.tensor(TOL_BF16))
    assert (diff <= step).all(), diff.max().item()
    assert (diff == 0).float().mean().item() >= 0.95


def test_page_ending_on_tile_boundary_followed_by_empty_pages(lis, oracle):
    """Regression (found by scripts/gpu_fuzz_pair.py): a page that ends exactly on a 256-row tile boundary, followed by
    two or more empty pages, with several query tiles resident -- the warps of the two column halves used to close the
    trailing empty page in different tiles, which mixed up the exchange slots of the single-CTA kernel."""
    from importlib import import_module

    N = import_module("multi-modal_colpali_b200._native")
    lib = N.load()
    g = torch.Generator().manual_seed(314)
    qs = ragged(g, [20] * 19 + [4])                       # 384 rows = 3 query tiles
    p_lens = [256, 0, 0, 100, 412, 0, 0, 0, 50, 206, 0, 0, 1024, 0, 0, 33]   # ends at rows 256, 768, 1024, 2048
    ps = ragged(g, p_lens)
    want = oracle.score_multi_vector_widened(qs, ps)
    try:
        for tun in ((0, 0, 1, 0, 1), (0, 2, 1, 0, 1), (128, 3, 1, 0, 1), (0, 0, 2, 0, 3), (0, 0, 0, 0, 0), (0, 0, 1, 0, 0)):
            N.check(lib.lis_set_tuning(*tun))
            got = lis.score_multi_vector(qs, ps, round_mode="f32")
            assert (got - want).abs().max().item() <= TOL_F32, tun
    finally:
        lib.lis_set_tuning(0, 0, 0, 0, 0)


def test_ragged_lists_and_zero_padding(lis, oracle):
    g = torch.Generator().manual_seed(3)
    q_lens = [5, 20, 33, 128, 130, 300, 1, 17]
    p_lens = [1, 2, 3, 31, 32, 33, 64, 127, 128, 129, 255, 256, 257, 300, 511, 513, 700, 768, 40, 7] * 9
    qs, ps = ragged(g, q_lens), ragged(g, p_lens)
    for bs in (128, 16, 7):
        check_both_modes(lis, oracle, qs, ps, batch_size=bs)


def test_negative_sims_are_clamped_like_zero_padding(lis, oracle):
    """A short page whose real tokens all point away from the query: the reference's pad rows give 0."""
    g = torch.Generator().manual_seed(4)
    q = rand_unit(g, 1, 8, 128)
    away = (-q[0, :3]).clone()                       # 3 tokens anti-aligned with query tokens 0..2
    long_page = rand_unit(g, 50, 128)
    ps = [away, long_page]
    want = oracle.score_multi_vector_widened(q, ps)
    got = lis.score_multi_vector(q, ps, round_mode="f32")
    assert (got - want).abs().max().item() <= TOL_F32
    # and without padding (a single page is its own longest page) the negative max survives
    want1 = oracle.score_multi_vector_widened(q, [away])
    got1 = lis.score_multi_vector(q, [away], round_mode="f32")
    assert (got1 - want1).abs().max().item() <= TOL_F32
    assert want1.item() < want[0, 0].item()


def test_padded_tensor_inputs_left_and_right(lis, oracle):
    """ColQwen pads on the left, ColPali on the right; encoders zero the padded rows."""
    g = torch.Generator().manual_seed(5)
    p = rand_unit(g, 70, 90, 128)
    for i in range(70):
        n_pad = int(torch.randint(0, 60, (1,), generator=g))
        if i % 2:
            p[i, :n_pad] = 0
        else:
            p[i, 90 - n_pad:] = 0
    q = rand_unit(g, 3, 24, 128)
    q[1, 20:] = 0
    check_both_modes(lis, oracle, q, p)


def test_known_answers(lis):
    """Analytic cases."""
    eye = torch.eye(128)
    q = eye[:10].to(torch.bfloat16).unsqueeze(0)                      # 10 one-hot query tokens
    pages = [eye[:64].to(torch.bfloat16), eye[5:69].to(torch.bfloat16), eye[64:].to(torch.bfloat16)]
    got = lis.score_multi_vector(q, pages, round_mode="f32")
    assert got.tolist() == [[10.0, 5.0, 0.0]]
    # appending a zero query token / permuting page tokens changes nothing
    q0 = torch.cat([q, torch.zeros(1, 1, 128, dtype=torch.bfloat16)], dim=1)
    perm = [p[torch.randperm(p.shape[0], generator=torch.Generator().manual_seed(1))] for p in pages]
    assert torch.equal(lis.score_multi_vector(q0, perm, round_mode="f32"), got)


def test_empty_inputs_raise(lis):
    x = torch.zeros(1, 4, 128, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="No queries provided"):
        lis.score_multi_vector([], x)
    with pytest.raises(ValueError, match="No passages provided"):
        lis.score_multi_vector(x, [])


def test_fp16_inputs(lis, oracle):
    g = torch.Generator().manual_seed(6)
    q = rand_unit(g, 4, 30, 128, dtype=torch.float16)
    p = rand_unit(g, 300, 100, 128, dtype=torch.float16)
    want = oracle.score_multi_vector_widened(q, p)
    got = lis.score_multi_vector(q, p, round_mode="f32")
    assert (got - want).abs().max().item() <= TOL_F32


def test_all_tilings_agree_bitwise(lis, oracle):
    """Every (tile_n, group) instantiation computes the same numbers."""
    from importlib import import_module

    N = import_module("multi-modal_colpali_b200._native")
    lib = N.load()
    g = torch.Generator().manual_seed(8)
    qs = ragged(g, [20] * 29 + [60])          # 640 rows -> 5 M tiles
    ps = ragged(g, [int(x) for x in torch.randint(1, 400, (500,), generator=g)])
    want = oracle.score_multi_vector_widened(qs, ps)
    base = base16 = None
    try:
        ss = [(256, 1), (256, 2), (256, 3), (128, 1), (128, 2), (128, 3), (128, 4), (128, 5)]
        ts = [(128, 1), (128, 2), (128, 3), (128, 4), (192, 1), (192, 2)]
        for a_op, tilings in ((1, ss), (2, ts)):
            for eh in (1, 2):
                for nt, grp in tilings:
                    N.check(lib.lis_set_tuning(nt, grp, 0, eh, a_op))
                    got = lis.score_multi_vector(qs, ps, round_mode="f32")
                    assert (got - want).abs().max().item() <= TOL_F32, (nt, grp, eh, a_op)
                    if base is None:
                        base = got
                    assert torch.equal(got, base), (nt, grp, eh, a_op)
                    got16 = lis.score_multi_vector(qs, ps)
                    if base16 is None:
                        base16 = got16
                    assert torch.equal(got16, base16), (nt, grp, eh, a_op)
        N.check(lib.lis_set_tuning(0, 0, 3, 0, 0))   # 3 CTAs only: long per-CTA page ranges
        got = lis.score_multi_vector(qs, ps, round_mode="f32")
        assert torch.equal(got, base)
    finally:
        lib.lis_set_tuning(0, 0, 0, 0, 0)


@pytest.mark.parametrize("n_tiles", [2, 3, 4, 5, 6, 7, 8, 9, 10, 13])
def test_pair_kernel_every_tile_count(lis, oracle, n_tiles):
    """CTA-pair form (cta_group::2): every pass shape -- M = 256 uses only (even tile counts) and a final
    64/64-split M = 128 use (odd) -- against the oracle, and bit-identical to the single-CTA form."""
    from importlib import import_module

    N = import_module("multi-modal_colpali_b200._native")
    lib = N.load()
    g = torch.Generator().manual_seed(100 + n_tiles)
    rows = n_tiles * 128 - 37                       # last tile partly filled
    pattern = [20, 33, 7, 64, 1, 150, 90, 128, 45]   # queries straddling 64- and 128-row boundaries
    q_lens, left, i = [], rows, 0
    while left > 0:
        n = min(pattern[i % len(pattern)], left)
        q_lens.append(n); left -= n; i += 1
    assert sum(q_lens) == rows
    qs = ragged(g, q_lens)
    p_lens = [int(x) for x in torch.randint(1, 700, (420,), generator=g)] + [0, 0, 1, 256, 512, 0, 31, 1030, 0]
    ps = ragged(g, p_lens)
    try:
        for bs in (128, 16):
            want = oracle.score_multi_vector_widened(qs, ps, batch_size=bs)
            N.check(lib.lis_set_tuning(0, 0, 0, 0, 1))            # single CTA per SM (SS form)
            single = lis.score_multi_vector(qs, ps, batch_size=bs, round_mode="f32")
            single16 = lis.score_multi_vector(qs, ps, batch_size=bs)
            for grp, ctas in ((10, 0), (8, 0), (6, 0), (4, 0), (3, 0), (0, 6), (0, 0)):
                N.check(lib.lis_set_tuning(0, grp, ctas, 0, 3))   # CTA pairs
                got = lis.score_multi_vector(qs, ps, batch_size=bs, round_mode="f32")
                assert (got - want).abs().max().item() <= TOL_F32, (n_tiles, grp, ctas, bs)
                assert torch.equal(got, single), (n_tiles, grp, ctas, bs)
                got16 = lis.score_multi_vector(qs, ps, batch_size=bs)
                assert torch.equal(got16, single16), (n_tiles, grp, ctas, bs)
            N.check(lib.lis_set_tuning(0, 0, 0, 0, 0))            # auto: the pass planner mixes both forms
            got = lis.score_multi_vector(qs, ps, batch_size=bs, round_mode="f32")
            assert torch.equal(got, single), (n_tiles, "auto", bs)
    finally:
        lib.lis_set_tuning(0, 0, 0, 0, 0)


def test_pair_kernel_fp16_and_uniform_pages(lis, oracle):
    g = torch.Generator().manual_seed(21)
    q = rand_unit(g, 32, 20, 128, dtype=torch.float16)           # 640 rows: 5 tiles, BASELINE configs[1] query shape
    p = rand_unit(g, 333, 1030, 128, dtype=torch.float16)
    want = oracle.score_multi_vector_widened(q, p)
    got = lis.score_multi_vector(q, p, round_mode="f32")
    assert (got - want).abs().max().item() <= TOL_F32


# ---------------------------------------------------------------------------------------------
def test_topk_matches_oracle_with_ties(lis, oracle):
    g = torch.Generator().manual_seed(9)
    scores = torch.randn(5, 20000, generator=g)
    scores[:, 1000:1010] = scores[:, :10]              # exact ties
    scores[2, 77] = float("nan")
    want_v, want_i = oracle.topk(torch.nan_to_num(scores, nan=float("-inf")), 100)
    got_v, got_i = lis.topk_device(scores.cuda(), 100)
    assert torch.equal(got_i.cpu(), want_i)
    assert torch.equal(got_v.cpu(), want_v)
    # k larger than the row, and a tiny row
    got_v, got_i = lis.topk_device(scores[:, :7].contiguous().cuda(), 10)
    assert (got_i[:, 7:] == -1).all() and torch.isinf(got_v[:, 7:]).all()
    want_v, want_i = oracle.topk(torch.nan_to_num(scores[:, :7], nan=float("-inf")), 7)
    assert torch.equal(got_i[:, :7].cpu(), want_i)


def test_topk_large_row_multi_pass(lis, oracle):
    g = torch.Generator().manual_seed(10)
    scores = torch.randn(2, 300_000, generator=g)
    for k in (10, 1024):
        want_v, want_i = oracle.topk(scores, k)
        got_v, got_i = lis.topk_device(scores.cuda(), k, id_base=5_000_000_000)
        assert torch.equal(got_i.cpu() - 5_000_000_000, want_i)
        assert torch.equal(got_v.cpu(), want_v)


def test_merge_topk(lis, oracle):
    g = torch.Generator().manual_seed(11)
    parts = []
    for r in range(4):
        v = torch.randn(3, 10, generator=g).sort(dim=1, descending=True).values
        i = torch.randint(0, 1 << 40, (3, 10), generator=g)
        if r == 3:
            v[:, 6:] = float("-inf"); i[:, 6:] = -1      # a short shard pads with -1
        parts.append((v, i))
    want_v, want_i = oracle.merge_topk(parts, 10)
    got_v, got_i = lis.merge_topk_device(torch.cat([p[0] for p in parts], 1).cuda(),
                                         torch.cat([p[1] for p in parts], 1).cuda(), 10)
    assert torch.equal(got_i.cpu(), want_i) and torch.equal(got_v.cpu(), want_v)


# ---------------------------------------------------------------------------------------------
def test_index_search_matches_oracle(lis, oracle):
    g = torch.Generator().manual_seed(12)
    p_lens = [int(x) for x in torch.randint(200, 769, (400,), generator=g)]
    ps = ragged(g, p_lens)
    qs = ragged(g, [32, 16, 45])
    idx = lis.LateInteractionIndex(sum(p_lens), len(ps))
    ids = idx.add(ps[:150], ids=list(range(1000, 1150)))
    idx.add(ps[150:], ids=list(range(5000, 5250)))
    assert len(idx) == 400 and idx.num_rows == sum(p_lens)
    all_ids = torch.tensor(list(range(1000, 1150)) + list(range(5000, 5250)))
    want = oracle.score_multi_vector_widened(qs, [p for p in ps], batch_size=10 ** 9)
    # an index built without zero_pad_block has mask (no-clamp) semantics; with unit-norm random
    # pages every query token has a positive best match, so both semantics coincide here
    full = idx.scores(qs).cpu()
    assert (full - want).abs().max().item() <= TOL_F32
    want_v, want_i = oracle.topk(full, 10)
    got_v, got_i = idx.search(qs, 10)
    assert torch.equal(got_i, all_ids[want_i]) and torch.equal(got_v, want_v)
    idx.close()


def test_index_synthetic_fill_and_planted_needles(lis, oracle):
    idx = lis.LateInteractionIndex(20000 * 64 + 1000, 20010)
    idx.fill_synthetic(20000, 64, seed=2004, id_base=0)
    rows = idx.read_rows(0, 64 * 20)
    assert torch.allclose(rows.float().norm(dim=-1), torch.ones(64 * 20), atol=1e-2)
    assert rows.float().std().item() == pytest.approx(1 / math.sqrt(128), rel=0.1)
    g = torch.Generator().manual_seed(1004)
    q = rand_unit(g, 1, 16, 128)
    needles = [unit(0.9 * q[0].float() + 0.02 * torch.randn(16, 128, generator=g)).to(torch.bfloat16) for _ in range(5)]
    idx.add(needles, ids=[900001 + i for i in range(5)])
    v, i = idx.search(q, 10)
    assert sorted(i[0, :5].tolist()) == [900001 + j for j in range(5)]
    assert v[0, 4].item() > v[0, 5].item() + 1.0
    # oracle check on a slice: first 20 synthetic pages
    want = oracle.score_multi_vector_widened(q, rows.reshape(20, 64, 128))
    got = idx.scores(q).cpu()[:, :20]
    assert (got - want).abs().max().item() <= TOL_F32
    idx.close()


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hidden", [768, 2048])
def test_projection_head(lis, oracle, hidden):
    g = torch.Generator().manual_seed(13)
    h = torch.randn(3, 301, hidden, generator=g).to(torch.bfloat16)
    w = (torch.randn(128, hidden, generator=g) / math.sqrt(hidden)).to(torch.bfloat16)
    b = (0.1 * torch.randn(128, generator=g)).to(torch.bfloat16)
    mask = (torch.rand(3, 301, generator=g) > 0.2).long()
    want = oracle.project_normalize(h.float(), w.float(), b.float(), mask.float())
    got = lis.project_normalize(h.cuda(), w.cuda(), b.cuda(), mask.cuda(), round_mode="f32").cpu()
    assert got.dtype == torch.bfloat16 and got.shape == (3, 301, 128)
    assert (got.float() - want).abs().max().item() <= 4e-3      # bf16 output rounding of values <= 1
    assert (got[mask == 0] == 0).all()
    got_nb = lis.project_normalize(h.cuda(), w.cuda(), round_mode="f32").cpu()
    want_nb = oracle.project_normalize(h.float(), w.float(), None, None)
    assert (got_nb.float() - want_nb).abs().max().item() <= 4e-3


# ---------------------------------------------------------------------------------------------
class _FakeBatch(dict):
    def to(self, device):
        return _FakeBatch({k: v.to(device) for k, v in self.items()})


class _FakeProcessor:
    def __init__(self, table):
        self.table = table

    def process_queries(self, queries):
        return _FakeBatch(emb=torch.stack([self.table[q] for q in queries]))


class _FakeModel:
    device = torch.device("cuda", 0)

    def __call__(self, emb):
        return emb


def test_reference_api_shapes(lis, oracle):
    g = torch.Generator().manual_seed(14)
    pages = rand_unit(g, 40, 50, 128)
    dataset = [{"embedding": pages[i], "doc_id": i // 4, "page_id": i % 4, "file_name": f"f{i // 4}.pdf"} for i in range(40)]
    images = {f"f{d}.pdf": {p: f"img-{d}-{p}" for p in range(4)} for d in range(10)}
    table = {"q0": rand_unit(g, 12, 128), "q1": rand_unit(g, 12, 128)}
    out = lis.score_results(["q0", "q1"], _FakeProcessor(table), _FakeModel(), dataset, images, top_k=5)
    want = oracle.score_multi_vector(torch.stack([table["q0"], table["q1"]]), pages)
    assert [len(r) for r in out] == [5, 5]
    for q in range(2):
        ids = [r["doc_id"] * 4 + r["page_id"] for r in out[q]]
        assert_topk_equiv(ids, [r["score"] for r in out[q]], want[q], TOL_BF16)
        assert all(r["image"] == images[r["file_name"]][r["page_id"]] for r in out[q])
        assert all(r["file_name"] == f"f{r['doc_id']}.pdf" for r in out[q])

    client = lis.MaxSimClient(capacity_rows=4096, capacity_pages=64)
    lis.ensure_colpali_collection(client, "colpali")
    pts = [lis.PointStruct(id=f"p{i}", vector=pages[i].float().tolist(),
                           payload={"document_name": f"f{i // 4}.pdf", "page_no": i % 4, "img_link": f"l{i}",
                                    "username": "ann" if i % 2 else "bob"}) for i in range(40)]
    client.upsert("colpali", pts[:25])
    client.upsert("colpali", pts[25:])
    assert client.count("colpali") == 40
    want32 = oracle.score_multi_vector_widened(table["q0"][None], pages)[0]
    res = lis.retrieve_colpali("q0", _FakeProcessor(table), _FakeModel(), client, "", "colpali", 5)
    ids = [int(p.id[1:]) for p in res.points]
    assert_topk_equiv(ids, [p.score for p in res.points], want32, 2e-2)   # cosine re-normalisation in bf16
    assert all(p.payload["img_link"] == f"l{i}" for p, i in zip(res.points, ids))
    res = lis.retrieve_colpali("q0", _FakeProcessor(table), _FakeModel(), client, "ann", "colpali", 5)
    odd = want32.clone(); odd[0::2] = float("-inf")
    ids = [int(p.id[1:]) for p in res.points]
    assert len(ids) == 5 and all(i % 2 == 1 for i in ids)
    assert_topk_equiv(ids, [p.score for p in res.points], odd, 2e-2)
    assert all(p.payload["username"] == "ann" for p in res.points)


# ---------------------------------------------------------------------------------------------
def test_fp32_embeddings_split_planes(lis, oracle):
    """ColFlor's default dtype (05_experiment02.py:343-347): fp32 in, fp32-accurate scores out."""
    g = torch.Generator().manual_seed(15)
    qs = [unit(torch.randn(n, 128, generator=g)) for n in (16, 20, 150)]
    p_lens = [int(x) for x in torch.randint(1, 300, (300,), generator=g)]
    ps = [unit(torch.randn(n, 128, generator=g)) for n in p_lens]
    want = oracle.score_multi_vector(qs, ps)                      # the reference's fp32 path
    got = lis.score_multi_vector(qs, ps)
    assert got.dtype == torch.float32 and got.shape == want.shape
    err = (got - want).abs().max().item()
    assert err <= TOL_F32, err
    rel = err / want.abs().max().item()
    assert rel <= 2e-6, f"split-fp32 should be fp32-accurate (relative), got {rel}"
    # padded tensor form + index form
    pt = unit(torch.randn(64, 77, 128, generator=g))
    qt = unit(torch.randn(5, 32, 128, generator=g))
    want = oracle.score_multi_vector(qt, pt)
    assert (lis.score_multi_vector(qt, pt) - want).abs().max().item() <= 2e-6 * want.abs().max().item()
    idx = lis.LateInteractionIndex(64 * 77, 64, dtype=torch.float32)
    idx.add(pt)
    assert torch.allclose(idx.read_rows(0, 77), pt[0], atol=1e-6)
    wv, wi = oracle.topk(want, 7)
    v, i = idx.search(qt, 7)
    assert torch.equal(i, wi) and (v - wv).abs().max().item() <= 2e-6 * want.abs().max().item()
    idx.close()


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_index_save_load_roundtrip(lis, oracle, tmp_path, dtype):
    g = torch.Generator().manual_seed(16)
    ps = [unit(torch.randn(int(n), 128, generator=g)).to(dtype) for n in torch.randint(1, 90, (120,), generator=g)]
    qs = [unit(torch.randn(12, 128, generator=g)).to(dtype), unit(torch.randn(31, 128, generator=g)).to(dtype)]
    idx = lis.LateInteractionIndex(sum(p.shape[0] for p in ps), len(ps), dtype=dtype)
    idx.add(ps, ids=[7 * i + 3 for i in range(len(ps))], payloads=[{"page_no": i} for i in range(len(ps))],
            zero_pad_block=128)
    v0, i0 = idx.search(qs, 9)
    idx.save(tmp_path / "ix")
    idx.close()
    back = lis.LateInteractionIndex.load(tmp_path / "ix", capacity_rows=20000, capacity_pages=200)
    assert len(back) == len(ps) and back.dtype == dtype and back.payloads[7 * 5 + 3] == {"page_no": 5}
    v1, i1 = back.search(qs, 9)
    assert torch.equal(i0, i1) and torch.equal(v0, v1)
    back.add(ps[:3], ids=[9001, 9002, 9003])              # still appendable after a load
    assert len(back) == len(ps) + 3
    back.close()


class _ImgProcessor:
    """process_images: a "page image" here is just an int seed; returns a dict batch like a HF processor."""

    def process_images(self, images):
        return _FakeBatch(seeds=torch.tensor(images))


class _ImgModel:
    device = torch.device("cuda", 0)

    def __call__(self, seeds):
        out = [unit(torch.randn(40, 128, generator=torch.Generator().manual_seed(int(s)))) for s in seeds.tolist()]
        return torch.stack(out).to(torch.bfloat16).cuda()


def test_ingestion_dropins(lis, oracle, tmp_path):
    import pickle

    model, proc = _ImgModel(), _ImgProcessor()
    images_per_pdf = {"a.pdf": [1, 2, 3], "b.pdf": [4, 5]}
    ds = lis.create_document_embeddings(images_per_pdf, model, proc, batch_size=2)
    assert [(e["doc_id"], e["page_id"], e["file_name"]) for e in ds] == [
        (0, 0, "a.pdf"), (0, 1, "a.pdf"), (0, 2, "a.pdf"), (1, 0, "b.pdf"), (1, 1, "b.pdf")]
    assert all(e["embedding"].device.type == "cpu" and e["embedding"].shape == (40, 128) for e in ds)
    with open(tmp_path / "emb.pkl", "wb") as f:           # the reference's cache format (05_experiment02.py:391-398)
        pickle.dump(ds, f)
    ds2 = lis.load_embedding_cache(str(tmp_path / "emb.pkl"))
    table = {"q": unit(torch.randn(10, 128, generator=torch.Generator().manual_seed(9))).to(torch.bfloat16)}
    res = lis.score_results(["q"], _FakeProcessor(table), _FakeModel(), ds2, {"a.pdf": "ABC", "b.pdf": "DE"}, top_k=3)
    want = oracle.score_multi_vector(table["q"][None], torch.stack([e["embedding"] for e in ds]))[0]
    assert_topk_equiv([{"a.pdf": 0, "b.pdf": 3}[r["file_name"]] + r["page_id"] for r in res[0]],
                      [r["score"] for r in res[0]], want, TOL_BF16)

    client = lis.MaxSimClient(capacity_rows=4096, capacity_pages=64)
    lis.ensure_colpali_collection(client, "pages")
    pages = [{"image": s, "filename": "a.pdf" if s <= 3 else "b.pdf", "page_no": s, "img_link": f"img{s}"} for s in range(1, 6)]
    lis.colpali_qdrant(pages, ["x/a.pdf", "y/b.pdf"], ["doi:a", "doi:b"], model, proc, client, "pages", batch_size=2)
    assert client.count("pages") == 5
    out = client.query_points("pages", query=table["q"], limit=5)
    assert len(out.points) == 5 and out.points[0].payload["type"] == "pdf_page"
    assert {p.payload["document_link"] for p in out.points} == {"doi:a", "doi:b"}
    best = out.points[0]
    assert best.payload["page_no"] == int(torch.argmax(want).item()) + 1


# ---------------------------------------------------------------------------------------------
def test_randomized_shapes_against_oracle(lis, oracle):
    """Seeded sweep over ragged shapes, including 0/1-token pages, many tiny pages per tile, queries
    longer than an M tile and page counts around the CTA count (148) and the 128-page block size."""
    g = torch.Generator().manual_seed(2026)
    for trial in range(12):
        nq = int(torch.randint(1, 9, (1,), generator=g))
        q_lens = [int(x) for x in torch.randint(1, 70, (nq,), generator=g)]
        if trial % 4 == 0:
            q_lens[0] = int(torch.randint(129, 400, (1,), generator=g))
        npg = [1, 2, 127, 128, 129, 147, 148, 149, 300, 1000, 2500, 37][trial]
        hi = [40, 3, 300, 1100, 16][trial % 5]
        p_lens = [int(x) for x in torch.randint(0 if trial % 3 == 0 else 1, hi + 1, (npg,), generator=g)]
        if sum(p_lens) == 0:
            p_lens[0] = 5
        qs, ps = ragged(g, q_lens), ragged(g, p_lens)
        bs = [128, 128, 32, 5][trial % 4]
        blocks_ok = all(max(p_lens[j:j + bs]) > 0 for j in range(0, npg, bs))
        if not blocks_ok:          # the reference cannot take the max over a block of empty pages
            continue
        want = oracle.score_multi_vector_widened(qs, ps, batch_size=bs)
        got = lis.score_multi_vector(qs, ps, batch_size=bs, round_mode="f32")
        assert got.shape == want.shape
        assert (got - want).abs().max().item() <= TOL_F32, (trial, nq, npg, hi, bs)


def test_many_queries_small_corpus(lis, oracle):
    """256 queries x 32 tokens = 64 M tiles -> 22 passes over the store (multi-pass bookkeeping)."""
    g = torch.Generator().manual_seed(17)
    q = rand_unit(g, 256, 32, 128)
    p = rand_unit(g, 150, 200, 128)
    want = oracle.score_multi_vector_widened(q, p)
    got = lis.score_multi_vector(q, p, round_mode="f32")
    assert (got - want).abs().max().item() <= TOL_F32
    idx = lis.LateInteractionIndex(150 * 200, 150)
    idx.add(p)
    wv, wi = oracle.topk(got, 20)
    v, i = idx.search(q, 20)
    assert torch.equal(i, wi) and torch.equal(v, wv)
    idx.close()


# ---------------------------------------------------------------------------------------------
def test_error_paths_and_limits(lis, oracle):
    """Capacity / argument errors surface as Python exceptions with the C-side message; big k; out= buffer."""
    g = torch.Generator().manual_seed(18)
    idx = lis.LateInteractionIndex(100, 3)
    with pytest.raises(ValueError, match="row capacity"):
        idx.add([rand_unit(g, 101, 128)])
    idx.add([rand_unit(g, 30, 128), rand_unit(g, 30, 128), rand_unit(g, 30, 128)])
    with pytest.raises(ValueError, match="page capacity"):
        idx.add([rand_unit(g, 1, 128)])
    with pytest.raises(ValueError, match="k out of range|k=|out of range"):
        idx.search(rand_unit(g, 1, 8, 128), 5000)
    with pytest.raises(ValueError):
        idx.search(torch.zeros(1, 8, 64, dtype=torch.bfloat16), 2)       # wrong embedding width
    v, i = idx.search(rand_unit(g, 2, 8, 128), 8)                          # k > number of pages
    assert (i[:, 3:] == -1).all() and torch.isinf(v[:, 3:]).all() and sorted(i[0, :3].tolist()) == [0, 1, 2]
    idx.close()
    with pytest.raises(ValueError, match="mix dtypes"):
        lis.score_multi_vector([rand_unit(g, 4, 128), rand_unit(g, 4, 128).float()], [rand_unit(g, 4, 128)])
    with pytest.raises(ValueError, match="queries are"):
        lis.score_multi_vector(rand_unit(g, 1, 4, 128), rand_unit(g, 2, 4, 128).to(torch.float16))
    # k = 1024 (LIS_MAX_K) over a multi-pass row, and the out= host buffer
    q = rand_unit(g, 3, 10, 128)
    p = rand_unit(g, 9000, 12, 128)
    out = torch.empty(3, 9000, dtype=torch.float32).pin_memory()
    got = lis.score_multi_vector(q, p, round_mode="f32", out=out)
    assert got is out
    want = oracle.score_multi_vector_widened(q, p)
    assert (out - want).abs().max().item() <= TOL_F32
    wv, wi = oracle.topk(out, 1024)
    gv, gi = lis.topk_device(out.cuda(), 1024)
    assert torch.equal(gi.cpu(), wi) and torch.equal(gv.cpu(), wv)
    with pytest.raises(ValueError, match="out must be"):
        lis.score_multi_vector(q, p, out=torch.empty(3, 5))


# ---------------------------------------------------------------------------------------------
def test_padded_ingestion_drops_pad_rows_bit_identically(lis, oracle):
    """add_padded stores no pad rows yet scores exactly like the reference on the padded tensor
    (left padding = ColQwen, right padding = ColPali)."""
    g = torch.Generator().manual_seed(19)
    B, S = 24, 70
    emb = rand_unit(g, B, S, 128)
    mask = torch.ones(B, S, dtype=torch.long)
    for i in range(B):
        n_pad = int(torch.randint(0, 50, (1,), generator=g))
        if i % 2:
            mask[i, :n_pad] = 0            # left padding
        else:
            mask[i, S - n_pad:] = 0        # right padding
    mask[3] = 1                            # one page without padding: must NOT be clamped
    emb = emb * mask.unsqueeze(-1).to(emb.dtype)
    q = rand_unit(g, 4, 16, 128)
    q[2, :5] = -emb[3, :5]                 # make some per-token maxima negative
    want = oracle.score_multi_vector_widened(q, emb)            # the reference on the padded tensor
    idx = lis.LateInteractionIndex(B * S, B)
    idx.add_padded(emb.cuda(), mask.cuda())
    assert idx.num_rows == int(mask.sum())                       # no pad rows in HBM
    got = idx.scores(q).cpu()
    assert (got - want).abs().max().item() <= TOL_F32
    padded = lis.score_multi_vector(q, emb, round_mode="f32")    # streaming the zero rows instead
    assert torch.equal(got, padded)
    idx.close()
    # fused ingestion from hidden states
    H = 256
    hidden = torch.randn(B, S, H, generator=g).to(torch.bfloat16)
    w = (torch.randn(128, H, generator=g) / 16).to(torch.bfloat16)
    b = (0.05 * torch.randn(128, generator=g)).to(torch.bfloat16)
    idx = lis.LateInteractionIndex(B * S, B)
    idx.add_from_hidden(hidden.cuda(), w.cuda(), b.cuda(), mask.cuda())
    e_ref = lis.project_normalize(hidden.cuda(), w.cuda(), b.cuda(), mask.cuda()).cpu()
    want = oracle.score_multi_vector_widened(q, e_ref)
    assert (idx.scores(q).cpu() - want).abs().max().item() <= TOL_F32
    idx.close()


# ---------------------------------------------------------------------------------------------
def test_full_size_config2_properties(lis, oracle):
    """BASELINE configs[1] at full size (32 queries x 20 tokens vs 100 000 pages x 1030 tokens, 26 GB):
    size-independent properties + oracle spot checks on pages read back from the store."""
    pages, ptok = 100_000, 1030
    idx = lis.LateInteractionIndex(pages * ptok, pages)
    idx.fill_synthetic(pages, ptok, seed=2002)
    g = torch.Generator().manual_seed(1002)
    q = rand_unit(g, 32, 20, 128)
    s1 = idx.scores(q)
    s2 = idx.scores(q)
    assert torch.equal(s1, s2)                                           # deterministic
    assert torch.isfinite(s1).all() and s1.shape == (32, pages)
    perm = torch.randperm(20, generator=g)
    s3 = idx.scores(q[:, perm])                                          # sum over query tokens is order-free ...
    assert (s3 - s1).abs().max().item() <= 1e-4                          # ... up to fp32 summation order
    # oracle on randomly chosen pages (rows copied back from HBM), incl. the first and the last page
    pick = [0, pages - 1] + torch.randint(1, pages - 1, (30,), generator=g).tolist()
    sub = torch.stack([idx.read_rows(p * ptok, ptok) for p in pick])
    want = oracle.score_multi_vector_widened(q, sub)
    got = s1[:, pick].cpu()
    assert (got - want).abs().max().item() <= TOL_F32
    # top-k: sorted, consistent with the matrix, identical to the oracle's rule on the full rows
    v, i = idx.search(q, 10)
    assert (v[:, :-1] >= v[:, 1:]).all()
    assert torch.equal(v, torch.gather(s1.cpu(), 1, i))
    wv, wi = oracle.topk(s1.cpu(), 10)
    assert torch.equal(i, wi) and torch.equal(v, wv)
    # reference-rounding mode stays within one bf16 step of the fp32 result everywhere
    s16 = idx.scores(q, round_mode="reference")
    assert ((s16 - s1).abs() <= 2 * bf16_step(s1)).all()
    idx.close()
