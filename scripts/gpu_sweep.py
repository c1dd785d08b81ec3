"""Tiling / regime sweep on one B200: K1 alone, device-resident inputs, CUDA-event timing.
Writes one JSON object per line to gpurun_out/sweep.jsonl."""
import importlib
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
native = importlib.import_module("multi-modal_colpali_b200._native")
scoring = importlib.import_module("multi-modal_colpali_b200.scoring")
lib = native.load()
dev = torch.device("cuda", 0)
out = open(ROOT / "gpurun_out" / "sweep.jsonl", "a")


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def time_k1(pq, store, iters=5, warm=2):
    scores = torch.empty((pq.plan.nq, store.n_pages), dtype=torch.float32, device=dev)
    for _ in range(warm):
        scoring.maxsim_scores_device(pq, store, "f32", out=scores)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        scoring.maxsim_scores_device(pq, store, "f32", out=scores)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def report(name, nq, qtok, pages, ptok, tilings, ragged=None):
    g = torch.Generator().manual_seed(1)
    q = unit(torch.randn(nq, qtok, 128, generator=g)).to(torch.bfloat16).to(dev)
    pq = scoring.pack_queries(q, dev)
    rows = pages * ptok if ragged is None else int(sum(ragged))
    idx = lis.LateInteractionIndex(rows, pages, device=dev)
    idx.fill_synthetic(pages, ptok if ragged is None else ragged, seed=7)
    store = idx._as_store()
    flops = 2.0 * nq * qtok * 128 * rows
    byts = rows * 256.0
    for til in tilings:
        nt, grp = til[0], til[1]
        eh = til[2] if len(til) > 2 else 0
        aop = til[3] if len(til) > 3 else 0
        native.check(lib.lis_set_tuning(nt, grp, 0, eh, aop))
        best, mean = time_k1(pq, store)
        rec = {"case": name, "nq": nq, "qtok": qtok, "pages": pages, "ptok": ptok, "rows": rows, "tile_n": nt,
               "group": grp, "epi_halves": eh, "a_operand": aop, "ms_best": best, "ms_mean": mean, "tflops": flops / (best * 1e-3) / 1e12,
               "gbs": byts / (best * 1e-3) / 1e9, "pairs_per_s": nq * pages / (best * 1e-3)}
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
        out.flush()
    lib.lis_set_tuning(0, 0, 0, 0, 0)
    idx.close()
    del store, idx
    torch.cuda.empty_cache()


which = sys.argv[1:] or ["c2", "hbm", "c5", "c3"]
if "c2" in which:
    report("c2_32x20_vs_100kx1030", 32, 20, 100_000, 1030,
           [(0, 0), (256, 3, 2, 1), (128, 4, 2, 2), (128, 3, 2, 2), (128, 2, 2, 2), (192, 2, 2, 2), (128, 3, 1, 2), (0, 0)])
if "hbm" in which:
    report("single_query_16tok_vs_100kx1030", 1, 16, 100_000, 1030, [(0, 0), (256, 1, 2, 1), (128, 1, 2, 2), (192, 1, 2, 2)])
    report("4q_32tok_vs_100kx1030", 4, 32, 100_000, 1030, [(0, 0), (256, 1, 2, 1), (192, 1, 2, 2)])
    report("8q_32tok_vs_100kx1030", 8, 32, 100_000, 1030, [(0, 0), (256, 2, 2, 1), (192, 2, 2, 2), (128, 2, 2, 2)])
if "c5" in which:
    report("c5_slice_1024x32_vs_20kx1030", 1024, 32, 20_000, 1030,
           [(0, 0), (256, 3, 2, 1), (128, 4, 2, 2), (128, 3, 2, 2), (192, 2, 2, 2), (128, 2, 2, 2), (128, 4, 1, 2)])
if "c3" in which:
    lens = torch.randint(256, 769, (200_000,), generator=torch.Generator().manual_seed(3003)).tolist()
    report("c3_slice_1q32_vs_200k_ragged256-768", 1, 32, 200_000, 0, [(0, 0), (256, 1, 1, 1)], ragged=lens)
    report("c3_slice_32q32_vs_200k_ragged256-768", 32, 32, 200_000, 0, [(0, 0), (256, 3, 2, 1)], ragged=lens)

if "k3" in which:
    import math
    head = importlib.import_module("multi-modal_colpali_b200.head")
    for n_tok, hidden in [(4 * 1030, 2048), (64 * 1030, 2048), (64 * 1030, 768), (256 * 1030, 1536)]:
        g = torch.Generator().manual_seed(5)
        h = torch.randn(n_tok, hidden, generator=g).to(torch.bfloat16).to(dev)
        w = (torch.randn(128, hidden, generator=g) / math.sqrt(hidden)).to(torch.bfloat16).to(dev)
        b = (0.1 * torch.randn(128, generator=g)).to(torch.bfloat16).to(dev)
        m = torch.ones(n_tok, dtype=torch.int64, device=dev)
        for _ in range(3):
            head.project_normalize(h, w, b, m)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); head.project_normalize(h, w, b, m); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        # torch's own route for the same op (cuBLAS GEMM + 3 elementwise kernels), as the existing-Blackwell baseline
        def ref():
            x = torch.nn.functional.linear(h, w, b)
            x = x / x.norm(dim=-1, keepdim=True)
            return x * m.unsqueeze(-1)
        for _ in range(3):
            ref()
        torch.cuda.synchronize()
        tr = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ref(); e1.record(); torch.cuda.synchronize()
            tr.append(e0.elapsed_time(e1))
        byts = n_tok * hidden * 2.0 + n_tok * 256.0
        rec = {"case": "k3_project_normalize", "n_tok": n_tok, "hidden": hidden, "ms_best": min(ts), "ms_torch_best": min(tr),
               "gbs": byts / (min(ts) * 1e-3) / 1e9, "tflops": 2.0 * n_tok * hidden * 128 / (min(ts) * 1e-3) / 1e12}
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n"); out.flush()

if "f32" in which:
    # fp32 embeddings (ColFlor's default dtype, 05_experiment02.py:343-347): two bf16 planes per operand, three MMAs per tile,
    # one query tile per pass.  Algorithmic bytes: 512 B per page-token row per pass; FLOP: 3 x the bf16 path's.
    for name, nq, qtok, pages in [("f32_c0_1q16_vs_1000x1030", 1, 16, 1000), ("f32_1q16_vs_50kx1030", 1, 16, 50_000),
                                   ("f32_c1_32q20_vs_20kx1030", 32, 20, 20_000)]:
        g = torch.Generator().manual_seed(1)
        q = unit(torch.randn(nq, qtok, 128, generator=g)).to(dev)
        pq = scoring.pack_queries(q, dev)
        idx = lis.LateInteractionIndex(pages * 1030, pages, dtype=torch.float32, device=dev)
        idx.fill_synthetic(pages, 1030, seed=7)
        store = idx._as_store()
        best, mean = time_k1(pq, store, iters=8, warm=3)
        n_pass = pq.plan.n_mtiles
        rows = pages * 1030
        rec = {"case": name, "dtype": "float32 (2 bf16 planes)", "nq": nq, "qtok": qtok, "pages": pages, "passes": n_pass,
               "ms_best": best, "ms_mean": mean, "gbs": rows * 512.0 * n_pass / (best * 1e-3) / 1e9,
               "tflops_useful": 2.0 * nq * qtok * 128 * rows / (best * 1e-3) / 1e12,
               "tflops_issued": 3 * 2.0 * n_pass * 128 * 128 * rows / (best * 1e-3) / 1e12,
               "pairs_per_s": nq * pages / (best * 1e-3)}
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n"); out.flush()
        idx.close()
