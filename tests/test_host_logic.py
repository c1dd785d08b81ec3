"""Host-side logic of the drop-in surface that needs no GPU."""
import numpy as np
import pytest
import torch


def test_clamp_flags_follow_reference_batching(lis, oracle):
    """clamp flag == "pad_sequence added zero rows to this page inside its batch"."""
    g = torch.Generator().manual_seed(0)
    lens = [int(x) for x in torch.randint(1, 9, (23,), generator=g)]
    for bs in (128, 8, 5, 1):
        flags = lis.clamp_flags(lens, bs)
        for j in range(0, len(lens), bs):
            blk = lens[j:j + bs]
            assert flags[j:j + bs].tolist() == [int(n < max(blk)) for n in blk]
    # semantic check against the oracle: the clamp only matters when it changes the result
    q = torch.randn(1, 4, 128, generator=g)
    pages = [-q[0, :2], torch.randn(6, 128, generator=g)]
    padded = oracle.score_multi_vector(q, pages)[0, 0]
    alone = oracle.score_multi_vector(q, pages[:1])[0, 0]
    assert lis.clamp_flags([2, 6]).tolist() == [1, 0] and padded > alone


def test_no_cpu_fallback(lis):
    x = torch.zeros(1, 4, 128, dtype=torch.bfloat16)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lis.score_multi_vector(x, x)
    with pytest.raises(RuntimeError, match="CUDA only|no CPU fallback"):
        lis.score_multi_vector(x, x, device="cpu")
    with pytest.raises(RuntimeError):
        lis.project_normalize(torch.zeros(2, 64, dtype=torch.bfloat16), torch.zeros(128, 64, dtype=torch.bfloat16))


def test_empty_inputs_raise_before_touching_the_gpu(lis):
    x = torch.zeros(1, 4, 128, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="No queries provided"):
        lis.score_multi_vector([], x)
    with pytest.raises(ValueError, match="No passages provided"):
        lis.score_multi_vector(x, [])


def test_shard_ranges(lis):
    for n, w in [(10, 3), (7, 8), (100_000, 8), (0, 2)]:
        parts = [lis.shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        lis.shard_range(10, 3, 3)
    lens = [5, 5, 5, 5, 100, 1, 1, 1]
    parts = lis.balanced_shard_ranges(lens, 2)
    assert parts == [(0, 4), (4, 8)]
    rng = np.random.default_rng(0)
    lens = rng.integers(256, 769, size=10_000)
    parts = lis.balanced_shard_ranges(lens, 8)
    assert parts[0][0] == 0 and parts[-1][1] == len(lens)
    tok = [int(lens[a:b].sum()) for a, b in parts]
    assert max(tok) / min(tok) < 1.01


def test_filter_parsing():
    import importlib

    api = importlib.import_module("multi-modal_colpali_b200.reference_api")
    f = {"must": [{"key": "username", "match": {"value": "ann"}}]}
    assert api._filter_conditions(f) == [("username", "ann")]
    assert api._filter_conditions(None) == []

    class MV:  # qdrant-client style objects
        def __init__(self, value): self.value = value

    class FC:
        def __init__(self, key, match): self.key, self.match = key, match

    class F:
        def __init__(self, must): self.must = must

    assert api._filter_conditions(F([FC("username", MV("bob"))])) == [("username", "bob")]


def test_embedding_cache_loader_validates(lis, tmp_path):
    import pickle

    good = [{"embedding": torch.zeros(3, 128), "doc_id": 0, "page_id": 0, "file_name": "a.pdf"}]
    with open(tmp_path / "ok.pkl", "wb") as f:
        pickle.dump(good, f)
    assert lis.load_embedding_cache(str(tmp_path / "ok.pkl"))[0]["file_name"] == "a.pdf"
    with open(tmp_path / "bad.pkl", "wb") as f:
        pickle.dump({"not": "a list"}, f)
    with pytest.raises(ValueError):
        lis.load_embedding_cache(str(tmp_path / "bad.pkl"))


def test_query_batcher_coalesces_and_routes_results(lis):
    """Host logic only: a fake index records the batches it is given."""
    import threading

    calls = []

    class FakeIndex:
        def search(self, qs, k, round_mode):
            calls.append(len(qs))
            s = torch.stack([torch.arange(k, 0, -1, dtype=torch.float32) + float(q[0, 0]) for q in qs])
            i = torch.stack([torch.arange(k, dtype=torch.int64) + int(q[0, 0]) * 100 for q in qs])
            return s, i

    b = lis.QueryBatcher(FakeIndex(), max_rows=64, max_wait_ms=50)
    outs = {}

    def client(j):
        q = torch.full((16, 128), float(j))
        outs[j] = b.search(q, 3 + (j % 2))

    ts = [threading.Thread(target=client, args=(j,)) for j in range(8)]
    [t.start() for t in ts]; [t.join() for t in ts]
    b.close()
    assert b.served == 8 and sum(calls) == 8 and max(calls) <= 4      # 64 rows = 4 queries of 16 tokens per pass
    assert len(calls) < 8                                              # something was coalesced
    for j in range(8):
        s, i = outs[j]
        assert len(s) == 3 + (j % 2) and i[0].item() == j * 100 and s[0].item() - j in (3.0, 4.0)   # kmax of its batch

    class Broken:
        def search(self, qs, k, round_mode):
            raise RuntimeError("boom")

    b = lis.QueryBatcher(Broken())
    with pytest.raises(RuntimeError, match="boom"):
        b.search(torch.zeros(4, 128), 2)
    b.close()


def test_query_batcher_keeps_queries_off_the_64_row_cuts(lis):
    """Bit-identity with one-at-a-time search needs every query to be cut into the same segments as when it is alone:
    the batcher pads with zero rows so that no query crosses a multiple of 64 packed rows unless it starts on one."""
    seen = []

    class FakeIndex:
        dtype = torch.float32

        def search(self, qs, k, round_mode):
            seen.append([int(q.shape[0]) for q in qs])
            assert all(float(q.abs().sum()) == 0 for q in qs if q.shape[0] not in (40, 70, 16, 64))   # fillers are zero rows
            n = len(qs)
            return torch.arange(n * k, dtype=torch.float32).reshape(n, k), torch.arange(n * k).reshape(n, k)

    b = lis.QueryBatcher(FakeIndex(), max_rows=256, max_wait_ms=200)
    futs = [b.submit(torch.ones(n, 128), 2) for n in (40, 40, 70, 16, 64)]
    b.close()
    [f.result(timeout=10) for f in futs]
    flat = [n for batch in seen for n in batch]
    assert [n for n in flat if n in (40, 70, 16, 64)] == [40, 40, 70, 16, 64]          # order of arrival kept
    for batch in seen:
        row = 0
        for n in batch:
            if n in (40, 70, 16, 64):      # a real query: either it fits before the next cut or it starts on one
                assert row % 64 == 0 or row % 64 + n <= 64, (batch, row, n)
            row += n
        assert row <= 256
    assert lis.QueryBatcher._placed_rows(40, 40) == 104 and lis.QueryBatcher._placed_rows(64, 70) == 134
    assert lis.QueryBatcher._placed_rows(10, 54) == 64


def test_shard_assignment_and_manifest_errors(lis, tmp_path):
    assert lis.assign_shards([5, 5, 5, 5], 4) == [(0, 1), (1, 2), (2, 3), (3, 4)]          # one shard per rank
    parts = lis.assign_shards([10, 10, 10, 10, 10, 10, 10, 10], 2)
    assert parts == [(0, 4), (4, 8)]
    parts = lis.assign_shards([100, 1, 1, 1], 2)                                          # balanced by rows, contiguous
    assert parts[0][0] == 0 and parts[-1][1] == 4 and parts[0][1] == parts[1][0]
    with pytest.raises(ValueError, match="manifest"):
        lis.LateInteractionIndex.read_manifest(tmp_path)
    (tmp_path / "manifest.json").write_text('{"format": "something-else", "dim": 128}')
    with pytest.raises(ValueError, match="lis-index-v2"):
        lis.LateInteractionIndex.read_manifest(tmp_path)


def test_dataset_fingerprint_sees_in_place_edits():
    import importlib

    api = importlib.import_module("multi-modal_colpali_b200.reference_api")
    ds = [{"embedding": torch.zeros(3, 128)} for _ in range(7)]
    fp = api._dataset_fingerprint(ds)
    assert api._dataset_fingerprint(ds) == fp
    ds[3] = {"embedding": torch.zeros(3, 128)}              # same length, another tensor
    assert api._dataset_fingerprint(ds) != fp
    fp = api._dataset_fingerprint(ds)
    ds[3]["embedding"].add_(1)                               # same tensor, edited in place
    assert api._dataset_fingerprint(ds) != fp


def test_product_never_imports_the_oracle_and_has_no_cpu_path():
    """The oracle is test infrastructure: nothing under the product package may import or execute it, and bench.py's own
    arm refuses to run without a GPU instead of falling back."""
    import re
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    for f in (root / "multi-modal_colpali_b200").rglob("*.py"):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
        assert "maxsim_oracle" not in text, f
    for f in (root / "multi-modal_colpali_b200" / "csrc").glob("*"):
        assert "oracle" not in f.read_text().lower(), f
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, str(root / "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
        assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


def test_flatten_queries_fast_and_slow_paths_agree():
    """The one-shot search packs queries on the host; data already in shape must pass through untouched
    (no copy), everything else must come out as ONE contiguous [rows, 128] matrix of the index dtype."""
    import importlib

    index = importlib.import_module("multi-modal_colpali_b200.index")
    scoring = importlib.import_module("multi-modal_colpali_b200.scoring")
    g = torch.Generator().manual_seed(5)
    q3 = torch.randn(4, 20, 128, generator=g).to(torch.bfloat16)
    flat, lens = index._flatten_queries(q3, torch.bfloat16)
    assert lens == (20,) * 4 and flat.shape == (80, 128) and flat.data_ptr() == q3.data_ptr()      # aliased, not copied
    flat32, _ = index._flatten_queries(q3.float(), torch.bfloat16)                                # dtype conversion
    assert flat32.dtype == torch.bfloat16 and torch.equal(flat32, flat)
    strided = torch.randn(4, 40, 128, generator=g).to(torch.bfloat16)[:, ::2]                     # not contiguous
    flat_s, lens_s = index._flatten_queries(strided, torch.bfloat16)
    assert flat_s.is_contiguous() and lens_s == (20,) * 4 and torch.equal(flat_s, strided.reshape(80, 128))
    ragged = [q3[0], q3[1][:7], torch.empty(0, 128, dtype=torch.bfloat16), q3[2][3:]]
    flat_r, lens_r = index._flatten_queries(ragged, torch.bfloat16)
    assert lens_r == (20, 7, 0, 17) and torch.equal(flat_r, torch.cat(ragged))
    one, lens_1 = index._flatten_queries([q3[3]], torch.bfloat16)
    assert lens_1 == (20,) and one.data_ptr() == q3[3].data_ptr()
    assert index._flatten_queries(tuple(ragged), torch.float16)[0].dtype == torch.float16
    assert index._flatten_queries((t for t in ragged), torch.bfloat16)[1] == lens_r               # any iterable
    with pytest.raises(ValueError):
        index._flatten_queries(torch.zeros(2, 5, 64), torch.bfloat16)
    with pytest.raises(ValueError):
        index._flatten_queries([torch.zeros(5, 64)], torch.bfloat16)
    with pytest.raises(ValueError):
        index._flatten_queries(torch.zeros(5, 128), torch.bfloat16)                                # one query needs a list
    # the plan carries the host addresses the C call receives; seg_first only when K1's rows are not the queries
    for lens in [(20,) * 4, (20,) * 10, (0, 100), (7,), (64, 64), (130,)]:
        plan = scoring.plan_queries(lens)
        lo, hi, mt, first = plan.host_ptrs
        assert (lo, hi, mt) == (plan.seg_lo.ctypes.data, plan.seg_hi.ctypes.data, plan.mt_seg.ctypes.data)
        assert (first is None) == plan.direct and (plan.direct or first == plan.seg_first.ctypes.data)
        assert plan.n_rows == sum(lens)
        assert scoring.plan_queries(list(lens)) is plan and scoring.plan_queries(np.asarray(lens)) is plan
