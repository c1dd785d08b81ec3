"""GPU parity tests added in round 2: every claimed behaviour of the new entry points runs under `pytest -m gpu`
through the C-ABI and is compared with the CPU oracle (or with the committed golden vectors)."""
import math
import os
import subprocess
import sys
import threading
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden" / "maxsim_golden.pt"
TOL_F32 = 1e-4      # north_star: inputs widened to fp32


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def rand_unit(gen, *shape, dtype=torch.bfloat16):
    return unit(torch.randn(*shape, generator=gen)).to(dtype)


def bf16_step(x):
    """Spacing of bf16 at |x| (8 significant bits)."""
    return torch.pow(2.0, torch.floor(torch.log2(x.abs().clamp_min(1e-30))) - 7)


# ---------------------------------------------------------------------------------------------
# 1. the committed golden vectors (outputs of the installed HF port of the reference's arithmetic) through CUDA
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["padded", "ragged", "ragged_bs8", "negative"])
def test_golden_fixtures_through_cuda(lis, case):
    gold = torch.load(GOLDEN)[case]
    qs, ps, bs = gold["qs"], gold["ps"], gold["batch_size"]
    got32 = lis.score_multi_vector(qs, ps, batch_size=bs, round_mode="f32")
    assert got32.dtype == torch.float32 and got32.device.type == "cpu"
    assert (got32 - gold["scores_fp32"]).abs().max().item() <= TOL_F32
    got16 = lis.score_multi_vector(qs, ps, batch_size=bs)              # reference rounding (default)
    diff = (got16 - gold["scores_bf16"]).abs()
    assert (diff <= torch.maximum(bf16_step(gold["scores_bf16"]), torch.tensor(1e-2))).all()
    assert (diff == 0).float().mean().item() >= 0.95
    # the same inputs resident on the device take the HBM-store route instead of the host-streaming route
    to_dev = (lambda x: x.cuda()) if isinstance(ps, torch.Tensor) else (lambda x: [t.cuda() for t in x])
    got_dev = lis.score_multi_vector(qs, to_dev(ps), batch_size=bs, round_mode="f32")
    assert torch.equal(got_dev, got32)


def test_golden_head_through_cuda(lis):
    gold = torch.load(GOLDEN)["head"]          # fp32 statements of HF modeling_colpali.py:148-155
    h, w, b, mask = gold["hidden"], gold["weight"], gold["bias"], gold["mask"]
    got = lis.project_normalize(h.to(torch.bfloat16).cuda(), w.to(torch.bfloat16).cuda(), b.to(torch.bfloat16).cuda(),
                                mask.cuda(), round_mode="f32").cpu()
    # inputs rounded to bf16 (rel. 2^-9 each, averaged over 192 terms) + one bf16 rounding of values <= 1
    assert (got.float() - gold["embeddings"]).abs().max().item() <= 8e-3
    assert (got[mask == 0] == 0).all()


# ---------------------------------------------------------------------------------------------
# 2. K3 against the reference head computed IN bf16 (what the reference's model stores)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hidden", [768, 2048])
def test_projection_head_reference_rounding(lis, oracle, hidden):
    """HF modeling_colpali.py:148-155 in the model dtype: Linear -> bf16, norm -> bf16, quotient -> bf16.  K3's
    round_mode="reference" rounds at the same places; what remains is the accumulation order of the 768/2048-term
    dot products (fp32 either way), which can flip a value that lands on a rounding boundary by one bf16 step."""
    g = torch.Generator().manual_seed(23)
    h = torch.randn(2, 257, hidden, generator=g).to(torch.bfloat16)
    w = (torch.randn(128, hidden, generator=g) / math.sqrt(hidden)).to(torch.bfloat16)
    b = (0.1 * torch.randn(128, generator=g)).to(torch.bfloat16)
    mask = (torch.rand(2, 257, generator=g) > 0.2).long()
    want = oracle.project_normalize(h, w, b, mask)                       # computed in bf16, like the reference model
    assert want.dtype == torch.bfloat16
    got = lis.project_normalize(h.cuda(), w.cuda(), b.cuda(), mask.cuda(), round_mode="reference").cpu()
    assert got.dtype == torch.bfloat16 and got.shape == want.shape
    diff = (got.float() - want.float()).abs()
    one_ulp = bf16_step(want.float().abs().clamp_min(2.0 ** -10)) * 1.001
    assert (diff <= 2 * one_ulp).all(), diff.max().item()       # a flipped Linear output AND a flipped quotient at worst
    share = (diff == 0).float().mean().item()
    assert share >= 0.90, f"only {share:.3f} of the stored values are bit-identical to the reference head"
    assert (diff <= one_ulp).float().mean().item() >= 0.999
    assert (got[mask == 0] == 0).all()
    # same-device check against torch's own bf16 route of the reference statements (cuBLAS Linear, bf16 norm and divide)
    hd, wd, bd = h.cuda(), w.cuda(), b.cuda()
    t = torch.nn.functional.linear(hd, wd, bd)
    t = t / t.norm(dim=-1, keepdim=True)
    t = (t * mask.cuda().unsqueeze(-1)).cpu()
    d2 = (got.float() - t.float()).abs()
    assert (d2 <= 2 * one_ulp).all() and (d2 == 0).float().mean().item() >= 0.90
    # the accuracy mode is closer to the exact unit vector than the reference's rounding chain
    exact = oracle.project_normalize(h.double(), w.double(), b.double(), mask.double())
    got32 = lis.project_normalize(h.cuda(), w.cuda(), b.cuda(), mask.cuda(), round_mode="f32").cpu()
    assert (got32.double() - exact).abs().mean() <= (got.double() - exact).abs().mean()


# ---------------------------------------------------------------------------------------------
# 3. one-shot search (lis_index_search_sharded with comm = NULL): graph replay == eager == oracle
# ---------------------------------------------------------------------------------------------
def test_one_shot_search_replays_as_graph_and_matches_oracle(lis, oracle):
    g = torch.Generator().manual_seed(31)
    lens = torch.randint(5, 400, (700,), generator=g).tolist()
    pages = [rand_unit(g, n, 128) for n in lens]
    idx = lis.LateInteractionIndex(sum(lens), len(lens))
    idx.add(pages, zero_pad_block=10 ** 9)
    cases = [
        [rand_unit(g, 16, 128)],                                              # single query (direct)
        [rand_unit(g, n, 128) for n in (20,) * 10],                           # reference chunk of 10 (cut at 64 rows)
        [rand_unit(g, n, 128) for n in (100, 3, 0, 130)],                     # split queries and an EMPTY one
        rand_unit(g, 5, 32, 128),                                             # padded tensor
    ]
    for qs in cases:
        ql = list(qs) if not isinstance(qs, torch.Tensor) else list(qs)
        full = oracle.score_multi_vector_widened([q for q in ql], pages, batch_size=10 ** 9) if all(
            q.shape[0] > 0 for q in ql) else None
        if full is None:     # the oracle's pad_sequence gives an empty query zero rows -> score 0 everywhere
            nz = [q if q.shape[0] else torch.zeros(1, 128, dtype=q.dtype) for q in ql]
            full = oracle.score_multi_vector_widened(nz, pages, batch_size=10 ** 9)
        for k in (1, 10, 100):
            wv, wi = oracle.topk(full, k)
            before = idx.graph_stats()
            v0, i0 = idx.search(qs, k)                    # eager + capture
            v1, i1 = idx.search(qs, k)                    # replay
            v2, i2 = idx.search(qs, k)
            after = idx.graph_stats()
            assert after[1] == before[1] + 1 and after[2] == before[2] + 2, (before, after)
            assert torch.equal(v0, v1) and torch.equal(i0, i1) and torch.equal(v1, v2) and torch.equal(i1, i2)
            assert (v0 - wv).abs().max().item() <= TOL_F32
            same = i0 == wi
            if not same.all():     # only exact-tie / tolerance-tie swaps are acceptable
                assert ((v0 - wv).abs() <= TOL_F32).all()
                assert (torch.gather(full, 1, i0.clamp_min(0)) - wv).abs().max().item() <= TOL_F32
            # the device-tensor route (lis_index_search) is the same kernels without the graph
            vd, idd = idx.search_device(qs, k)
            assert torch.equal(vd.cpu(), v0) and torch.equal(idd.cpu(), i0)
    # queries that already live on the device take the D2D staging route
    qd = [rand_unit(g, 16, 128).cuda(), rand_unit(g, 20, 128).cuda()]
    vh, ih = idx.search([q.cpu() for q in qd], 7)
    vg, ig = idx.search(qd, 7)
    vg2, ig2 = idx.search(qd, 7)
    assert torch.equal(vh, vg) and torch.equal(ih, ig) and torch.equal(vg, vg2) and torch.equal(ig, ig2)
    # the corpus grew: cached graphs for the old size must not be replayed
    extra = [rand_unit(g, 33, 128) for _ in range(5)]
    extra[2] = torch.cat([qd[0].cpu(), extra[2]])
    idx2 = lis.LateInteractionIndex(sum(lens) + 400, len(lens) + 5)
    idx2.add(pages)
    a = idx2.search([qd[0].cpu()], 3)
    idx2.add(extra)
    b = idx2.search([qd[0].cpu()], 3)
    assert b[1][0, 0].item() == len(lens) + 2 and a[1][0, 0].item() != len(lens) + 2
    idx.close(); idx2.close()


def test_one_shot_search_fp32_index(lis, oracle):
    g = torch.Generator().manual_seed(32)
    pages = [unit(torch.randn(n, 128, generator=g)) for n in torch.randint(10, 120, (150,), generator=g).tolist()]
    qs = [unit(torch.randn(n, 128, generator=g)) for n in (16, 40)]
    idx = lis.LateInteractionIndex(sum(p.shape[0] for p in pages), len(pages), dtype=torch.float32)
    idx.add(pages)
    full = oracle.score_multi_vector(qs, pages, batch_size=10 ** 9)        # fp32 reference
    wv, wi = oracle.topk(full, 10)
    for _ in range(3):
        v, i = idx.search(qs, 10)
        assert torch.equal(i, wi) and (v - wv).abs().max().item() <= TOL_F32
    idx.close()


# ---------------------------------------------------------------------------------------------
# 4. QueryBatcher on a real index
# ---------------------------------------------------------------------------------------------
def test_query_batcher_on_real_index_bit_identical(lis, oracle):
    g = torch.Generator().manual_seed(41)
    lens = torch.randint(20, 300, (900,), generator=g).tolist()
    pages = [rand_unit(g, n, 128) for n in lens]
    idx = lis.LateInteractionIndex(sum(lens), len(lens))
    idx.add(pages)
    n_clients = 24
    queries = [rand_unit(g, int(n), 128) for n in torch.randint(8, 40, (n_clients,), generator=g)]
    ks = [1 + (7 * j) % 23 for j in range(n_clients)]
    alone = [idx.search([q], k) for q, k in zip(queries, ks)]          # one at a time
    full = oracle.score_multi_vector_widened(queries, pages, batch_size=10 ** 9)
    batcher = lis.QueryBatcher(idx, max_rows=256, max_wait_ms=20)
    outs = [None] * n_clients
    start = threading.Barrier(n_clients)

    def client(j):
        start.wait()
        outs[j] = batcher.search(queries[j], ks[j])

    ts = [threading.Thread(target=client, args=(j,)) for j in range(n_clients)]
    [t.start() for t in ts]; [t.join() for t in ts]
    # requests queued right before close() are still served
    late = [batcher.submit(queries[j], ks[j]) for j in range(4)]
    batcher.close()
    assert batcher.served == n_clients + 4 and batcher.batches < n_clients      # something was coalesced
    with pytest.raises(RuntimeError, match="closed"):
        batcher.submit(queries[0], 3)
    for j in range(n_clients):
        s, i = outs[j]
        assert torch.equal(s, alone[j][0][0]) and torch.equal(i, alone[j][1][0]), j     # bit-identical to one-at-a-time
        wv, wi = oracle.topk(full[j:j + 1], ks[j])
        assert torch.equal(i, wi[0]) and (s - wv[0]).abs().max().item() <= TOL_F32
    for j, fut in enumerate(late):
        s, i = fut.result(timeout=30)
        assert torch.equal(s, alone[j][0][0]) and torch.equal(i, alone[j][1][0])
    # a malformed request fails alone, at submit time
    b2 = lis.QueryBatcher(idx)
    with pytest.raises(ValueError):
        b2.submit(torch.zeros(4, 64), 3)
    with pytest.raises(ValueError):
        b2.submit(queries[0], 5000)
    assert torch.equal(b2.search(queries[0], ks[0])[1], alone[0][1][0])
    b2.close()
    idx.close()


# ---------------------------------------------------------------------------------------------
# 5. host-resident corpus through surface 1 (the reference's literal call: ps on the CPU)
# ---------------------------------------------------------------------------------------------
def test_host_resident_corpus_streams_in_chunks(lis, oracle):
    scoring = __import__("importlib").import_module("multi-modal_colpali_b200.scoring")
    g = torch.Generator().manual_seed(51)
    n_pages = 10_240
    lens = torch.randint(1, 64, (n_pages,), generator=g).tolist()
    lens[5] = 0
    lens[n_pages - 1] = 0
    pages = [rand_unit(g, n, 128) for n in lens]
    qs = [rand_unit(g, n, 128) for n in (16, 20, 70)]
    qs[1][:4] = -pages[9][:4] if lens[9] >= 4 else qs[1][:4]
    want = oracle.score_multi_vector_widened(qs, pages)                   # 128-page blocks, zero padding
    got = lis.score_multi_vector(qs, pages, round_mode="f32")            # list of CPU tensors -> lis_stream_scores
    assert got.shape == (3, n_pages) and (got - want).abs().max().item() <= TOL_F32
    ondev = lis.score_multi_vector(qs, [p.cuda() for p in pages], round_mode="f32")
    assert torch.equal(got, ondev)                                        # same bits as the HBM-resident route
    # small chunks (many chunk boundaries, both buffers reused many times), one thread and several
    dev = torch.device("cuda", torch.cuda.current_device())
    pq = scoring.pack_queries(qs, dev)
    for chunk_rows, threads in ((4096, 1), (10_000, 4), (1 << 19, 0)):
        out = scoring.stream_scores_host_corpus(pq, pages, 128, "f32", chunk_rows=chunk_rows, host_threads=threads).cpu()
        assert torch.equal(out, got), (chunk_rows, threads)
    # contiguous [n, S, 128] tensor, pageable and pinned (the pinned one is DMA'd in place)
    dense = rand_unit(g, 600, 50, 128)
    w2 = oracle.score_multi_vector_widened(qs, dense)
    for t in (dense, dense.pin_memory()):
        out = scoring.stream_scores_host_corpus(pq, t, 128, "f32", chunk_rows=5000).cpu()
        assert (out - w2).abs().max().item() <= TOL_F32
    ref16 = lis.score_multi_vector(qs, dense)                             # default reference rounding on the host route
    dev16 = lis.score_multi_vector(qs, dense.cuda())
    assert torch.equal(ref16, dev16)
    importlib = __import__("importlib")
    importlib.import_module("multi-modal_colpali_b200._native").load().lis_stream_release()


# ---------------------------------------------------------------------------------------------
# 6. ingestion fusion: K3 writes compacted rows straight into the page store
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("round_mode", ["reference", "f32"])
def test_fused_ingestion_compacts_in_the_kernel(lis, oracle, round_mode):
    g = torch.Generator().manual_seed(61)
    B, S, H = 37, 90, 320
    hidden = torch.randn(B, S, H, generator=g).to(torch.bfloat16)
    w = (torch.randn(128, H, generator=g) / math.sqrt(H)).to(torch.bfloat16)
    b = (0.05 * torch.randn(128, generator=g)).to(torch.bfloat16)
    mask = torch.ones(B, S, dtype=torch.long)
    for i in range(B):
        n_pad = int(torch.randint(0, 60, (1,), generator=g))
        if i % 3 == 0:
            mask[i, :n_pad] = 0                # left padding (ColQwen)
        elif i % 3 == 1:
            mask[i, S - n_pad:] = 0            # right padding (ColPali)
        else:
            mask[i, torch.randperm(S, generator=g)[:n_pad]] = 0      # holes anywhere
    mask[4] = 1
    mask[5] = 0                                # a page with no tokens at all
    idx = lis.LateInteractionIndex(B * S + 200, B + 3)
    first = [rand_unit(g, 11, 128), rand_unit(g, 23, 128)]
    idx.add(first)                             # the store is not empty: rows must land behind what is there
    ids = idx.add_from_hidden(hidden.cuda(), w.cuda(), b.cuda(), mask.cuda(), ids=list(range(100, 100 + B)),
                              round_mode=round_mode)
    assert ids.tolist() == list(range(100, 100 + B)) and len(idx) == B + 2
    assert idx.num_rows == 34 + int(mask.sum())                         # no pad rows in HBM
    assert idx.page_lens(2).tolist() == mask.sum(dim=1).tolist()
    # the rows in the store == the dense K3 output with the pad rows dropped, bit for bit
    dense = lis.project_normalize(hidden.cuda(), w.cuda(), b.cuda(), mask.cuda(), round_mode=round_mode).cpu()
    want_rows = dense[mask.bool()]
    got_rows = idx.read_rows(34, int(mask.sum()))
    assert torch.equal(got_rows, want_rows)
    assert torch.equal(idx.read_rows(0, 34), torch.cat(first))
    # clamp flags == "the reference padded this page" -> scores equal the reference on the padded tensor
    q = rand_unit(g, 3, 16, 128)
    q[1, :6] = -dense[4, :6]
    want = oracle.score_multi_vector_widened(q, dense)
    got = idx.scores(q).cpu()[:, 2:]
    assert (got - want).abs().max().item() <= TOL_F32
    # int32 / bool masks are taken as they are
    for m in (mask.to(torch.int32), mask.bool()):
        idx3 = lis.LateInteractionIndex(B * S, B)
        idx3.add_from_hidden(hidden.cuda(), w.cuda(), b.cuda(), m.cuda(), round_mode=round_mode)
        assert torch.equal(idx3.read_rows(0, idx3.num_rows), want_rows)
        idx3.close()
    with pytest.raises(ValueError, match="capacity"):
        small = lis.LateInteractionIndex(10, B)
        small.add_from_hidden(hidden.cuda(), w.cuda(), b.cuda(), mask.cuda())
    idx.close()


# ---------------------------------------------------------------------------------------------
# 7. multi-GPU: identical results for every world size the box offers -- driver-run
# ---------------------------------------------------------------------------------------------
def test_multi_gpu_world_invariance():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on the box (the 1-GPU form of the same path is test_one_shot_search_*)")
    for world in [w for w in (2, 4, 8) if w <= n]:
        port = 29600 + (os.getpid() + world) % 300
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "scripts" / "sharded_check.py")]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=str(ROOT))
        assert r.returncode == 0 and "sharded-check PASS" in r.stdout, (world, r.stdout[-2000:], r.stderr[-2000:])


def test_single_rank_communicator_is_a_noop(lis, oracle):
    """world = 1 through the communicator API: lis_comm_init without NCCL, search == plain search."""
    sharded = __import__("importlib").import_module("multi-modal_colpali_b200.sharded")
    g = torch.Generator().manual_seed(71)
    pages = [rand_unit(g, 40, 128) for _ in range(64)]
    idx = lis.LateInteractionIndex(64 * 40, 64)
    idx.add(pages)
    comm = sharded.Communicator(idx.device)
    assert comm.world == 1 and comm.rank == 0
    q = [rand_unit(g, 16, 128)]
    a = idx.search(q, 5)
    b = idx.search(q, 5, comm=comm.handle)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    sh = lis.ShardedIndex(idx)
    c = sh.search(q, 5)
    assert torch.equal(a[1], c[1])
    comm.close(); idx.close()


# ---------------------------------------------------------------------------------------------
# 8. Qdrant-shaped client: fp32 storage by default, growth instead of silent loss, upsert replaces by id
# ---------------------------------------------------------------------------------------------
def test_maxsim_client_fp32_growth_and_upsert_by_id(lis, oracle):
    g = torch.Generator().manual_seed(81)
    n = 90
    raw = [torch.randn(int(t), 128, generator=g) * 3.0 for t in torch.randint(5, 60, (n,), generator=g)]   # NOT unit norm
    client = lis.MaxSimClient(capacity_rows=256, capacity_pages=8)      # far too small: must grow, not drop
    lis.ensure_colpali_collection(client, "c")
    pts = [lis.PointStruct(id=f"p{i}", vector=raw[i].tolist(), payload={"page_no": i, "username": "ann" if i % 3 else "bob"})
           for i in range(n)]
    for a in range(0, n, 7):
        client.upsert("c", pts[a:a + 7])
    assert client.count("c") == n
    cos = [unit(p) for p in raw]                                        # what Distance.COSINE stores
    q = torch.randn(19, 128, generator=g)

    def qdrant_maxsim(qv, pages):
        """Qdrant's MAX_SIM comparator on the stored (cosine-normalised) multivectors: sum over query vectors of the
        best dot product among the page's OWN vectors -- no padding rows, unlike score_multi_vector's batches."""
        return torch.stack([(unit(qv) @ pg.T).max(dim=1).values.sum() for pg in pages])

    want = qdrant_maxsim(q, cos)
    res = client.query_points("c", q.tolist(), limit=10)
    wv, wi = oracle.topk(want[None], 10)
    assert [int(p.id[1:]) for p in res.points] == wi[0].tolist()
    assert max(abs(p.score - v) for p, v in zip(res.points, wv[0].tolist())) <= TOL_F32     # fp32 planes: no 1e-2 slack
    # replace a point: the old version must disappear, the count must not change
    best = wi[0, 0].item()
    newvec = torch.randn(33, 128, generator=g)
    client.upsert("c", [lis.PointStruct(id=f"p{best}", vector=newvec.tolist(), payload={"page_no": best, "v": 2})])
    assert client.count("c") == n
    cos2 = list(cos)
    cos2[best] = unit(newvec)
    want2 = qdrant_maxsim(q, cos2)
    wv2, wi2 = oracle.topk(want2[None], n)
    res2 = client.query_points("c", q.tolist(), limit=n)
    ids2 = [int(p.id[1:]) for p in res2.points]
    assert ids2 == wi2[0].tolist() and len(set(ids2)) == n              # one entry per point id: the old version is gone
    assert max(abs(p.score - v) for p, v in zip(res2.points, wv2[0].tolist())) <= TOL_F32
    assert [p.payload.get("v") for p in res2.points if p.id == f"p{best}"] == [2]
    # filters only see current versions
    res3 = client.query_points("c", q.tolist(), limit=n,
                               query_filter={"must": [{"key": "username", "match": {"value": "bob"}}]})
    assert sorted(int(p.id[1:]) for p in res3.points) == [i for i in range(n) if i % 3 == 0 and i != best]
    client.delete_collection("c")


def test_dataset_index_cache_is_bounded_and_content_checked(lis, oracle):
    api = __import__("importlib").import_module("multi-modal_colpali_b200.reference_api")
    g = torch.Generator().manual_seed(82)
    lis.invalidate_dataset_index()
    mk = lambda: [{"embedding": rand_unit(g, 20, 128), "doc_id": 0, "page_id": i, "file_name": "f"} for i in range(12)]
    ds = mk()
    a = lis.index_for_dataset(ds)
    assert lis.index_for_dataset(ds) is a                               # cached
    ds[5] = {"embedding": rand_unit(g, 20, 128), "doc_id": 0, "page_id": 5, "file_name": "f"}   # same length, new content
    b = lis.index_for_dataset(ds)
    assert b is not a
    q = rand_unit(g, 1, 16, 128)
    want = oracle.score_multi_vector_widened(q, [e["embedding"] for e in ds])
    assert (b.scores(q).cpu() - want).abs().max().item() <= TOL_F32
    others = [mk() for _ in range(api._DATASET_CACHE_MAX + 2)]
    for o in others:
        lis.index_for_dataset(o)
    assert len(api._DATASET_INDEX) == api._DATASET_CACHE_MAX           # bounded: the oldest were closed
    lis.invalidate_dataset_index()
    assert len(api._DATASET_INDEX) == 0


# ---------------------------------------------------------------------------------------------
# 9. sharded on-disk format: cut once, load on any world size; pickle-cache converter
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_sharded_directory_roundtrip(lis, oracle, tmp_path, dtype):
    import json
    import pickle

    g = torch.Generator().manual_seed(91)
    ps = [unit(torch.randn(int(t), 128, generator=g)).to(dtype) for t in torch.randint(1, 200, (333,), generator=g)]
    ids = [5 * i + 2 for i in range(len(ps))]
    idx = lis.LateInteractionIndex(sum(p.shape[0] for p in ps), len(ps), dtype=dtype)
    idx.add(ps, ids=ids, payloads=[{"page_no": i} for i in range(len(ps))], zero_pad_block=128)
    qs = [unit(torch.randn(t, 128, generator=g)).to(dtype) for t in (16, 33)]
    v0, i0 = idx.search(qs, 12)
    man = idx.save(tmp_path / "ix", shards=4)
    assert man["format"] == "lis-index-v2" and len(man["shards"]) == 4
    assert sum(s["n_pages"] for s in man["shards"]) == len(ps) and sum(s["n_rows"] for s in man["shards"]) == idx.num_rows
    rows = [s["n_rows"] for s in man["shards"]]
    assert max(rows) / max(min(rows), 1) < 1.3                                    # cut by token count
    assert json.loads((tmp_path / "ix" / "manifest.json").read_text())["n_pages"] == len(ps)
    whole = lis.LateInteractionIndex.load(tmp_path / "ix")
    v1, i1 = whole.search(qs, 12)
    assert torch.equal(i0, i1) and torch.equal(v0, v1) and whole.payloads[ids[7]] == {"page_no": 7}
    assert torch.equal(whole.read_rows(0, whole.num_rows), idx.read_rows(0, idx.num_rows))
    # "world = 2" by hand: each half loads its shards; merged top-k == unsharded top-k
    halves = lis.assign_shards(rows, 2)
    parts = []
    for a, b in halves:
        part = lis.LateInteractionIndex.load(tmp_path / "ix", shard_ids=list(range(a, b)))
        parts.append(part.search(qs, 12))
        part.close()
    mv, mi = oracle.merge_topk(parts, 12)
    assert torch.equal(mi, i0) and torch.equal(mv, v0)
    whole.close()
    # pickle payloads are refused unless the caller vouches for the directory
    idx.payloads[ids[0]] = {"obj": object()}
    idx.save(tmp_path / "pk")
    with pytest.raises(ValueError, match="allow_pickle"):
        lis.LateInteractionIndex.load(tmp_path / "pk")
    lis.LateInteractionIndex.load(tmp_path / "pk", allow_pickle=True).close()
    idx.close()
    # converter from the reference's pickle cache (05_experiment02.py:391-398)
    if dtype == torch.bfloat16:
        dataset = [{"embedding": ps[i], "doc_id": i // 3, "page_id": i % 3, "file_name": f"d{i // 3}.pdf"} for i in range(60)]
        with open(tmp_path / "cache.pkl", "wb") as f:
            pickle.dump(dataset, f)
        man = lis.convert_embedding_cache(str(tmp_path / "cache.pkl"), str(tmp_path / "conv"), shards=3)
        assert len(man["shards"]) == 3 and man["n_pages"] == 60
        conv = lis.LateInteractionIndex.load(tmp_path / "conv")
        direct = lis.index_for_dataset(dataset)
        a, b = conv.search(qs, 5, round_mode="reference"), direct.search(qs, 5, round_mode="reference")
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
        assert conv.payloads[4] == {"doc_id": 1, "page_id": 1, "file_name": "d1.pdf"}
        conv.close()
        lis.invalidate_dataset_index()


# ---------------------------------------------------------------------------------------------
# 10. large results leave the device in column blocks behind the kernel (the D2H leg of surface 1)
# ---------------------------------------------------------------------------------------------
def test_large_result_is_copied_out_in_blocks_behind_the_kernel(lis, oracle):
    g = torch.Generator().manual_seed(101)
    n_pages = 9001                                                      # >= 8192: the overlapped route
    pages = rand_unit(g, n_pages, 9, 128).cuda()
    qs = [rand_unit(g, n, 128) for n in (16, 70, 3)]                    # one query is cut at a 64-row boundary
    want = oracle.score_multi_vector_widened(qs, pages.cpu())
    got = lis.score_multi_vector(qs, pages, round_mode="f32")
    assert got.shape == (3, n_pages) and got.device.type == "cpu" and (got - want).abs().max().item() <= TOL_F32
    dev = lis.score_multi_vector(qs, pages, round_mode="f32", return_device=True)       # the plain route
    assert torch.equal(dev.cpu(), got)
    pinned = torch.empty((3, n_pages), dtype=torch.float32).pin_memory()
    assert lis.score_multi_vector(qs, pages, round_mode="f32", out=pinned) is pinned and torch.equal(pinned, got)
    ref16 = lis.score_multi_vector(qs, pages)                           # reference rounding through the same route
    assert torch.equal(ref16, lis.score_multi_vector(qs, pages, return_device=True).cpu())
    with pytest.raises(ValueError, match="out must be"):
        lis.score_multi_vector(qs, pages, out=torch.empty(3, 5))


def test_pass_cost_calibration_installs_a_table(lis, oracle):
    """calibrate_pass_costs measures this device and installs the table; scoring stays correct under it."""
    N = __import__("importlib").import_module("multi-modal_colpali_b200._native")
    try:
        costs = lis.calibrate_pass_costs(pages=1500, iters=2)
        assert all(c > 0 for c in costs["single"][1:]) and all(c > 0 for c in costs["pair"][2:])
        assert costs["pair"][10] > costs["pair"][2]
        g = torch.Generator().manual_seed(111)
        q = rand_unit(g, 40, 32, 128)                                   # 10 query tiles
        pages = [rand_unit(g, n, 128) for n in (300, 17, 1030, 64) * 6]
        want = oracle.score_multi_vector_widened(q, pages)
        got = lis.score_multi_vector(q, [p.cuda() for p in pages], round_mode="f32")
        assert (got - want).abs().max().item() <= TOL_F32
    finally:
        N.check(N.load().lis_set_pass_costs(None, None))
