"""Full-size single-GPU shards of BASELINE configs 3/4/5 (the per-GPU work of the 8-GPU configs):
  c4: 500 000 pages x 1030 tokens (131.8 GB), 1 query x 16 tokens, top-10 -> p50 latency vs HBM floor
  c3: 1 000 000 ragged pages (256..768 tokens, 131 GB), 1x32 and 32x32 queries, top-100
  c5: 1024 queries x 32 tokens vs 200 000 pages x 1030 (52.7 GB): full [1024, 200k] score matrix
Writes JSON lines to gpurun_out/scale.jsonl."""
import importlib
import json
import statistics
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
scoring = importlib.import_module("multi-modal_colpali_b200.scoring")
dev = torch.device("cuda", 0)
out = open(ROOT / "gpurun_out" / "scale.jsonl", "a")
HBM = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
TF = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["bf16_tflops"] if (ROOT / "MEASURED_PEAKS.json").exists() else 1590.0


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def emit(rec):
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")
    out.flush()


def queries(nq, ntok, seed):
    return unit(torch.randn(nq, ntok, 128, generator=torch.Generator().manual_seed(seed))).to(torch.bfloat16)


def latency(idx, q_host, k, iters, warm=10):
    for _ in range(warm):
        idx.search(q_host, k)
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        idx.search(q_host, k)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[int(len(ts) * 0.95)]


def device_ms(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), statistics.mean(ts)


which = sys.argv[1:] or ["c4", "c3", "c5"]

if "c4" in which:
    pages, ptok = 500_000, 1030
    idx = lis.LateInteractionIndex(pages * ptok + 64 * 16, pages + 16, device=dev)
    t0 = time.perf_counter()
    idx.fill_synthetic(pages, ptok, seed=2004, id_base=0)
    torch.cuda.synchronize()
    fill_s = time.perf_counter() - t0
    q = queries(1, 16, 1004)
    # plant 10 needles with known ids so the top-10 is known without an oracle pass over 132 GB
    g = torch.Generator().manual_seed(5)
    needles = [unit(0.9 * q[0].float() + 0.02 * torch.randn(16, 128, generator=g)).to(torch.bfloat16) for _ in range(10)]
    idx.add(needles, ids=[10_000_000 + i for i in range(10)])
    qp = q.pin_memory()
    v, i = idx.search(qp, 10)
    needles_found = sorted(i[0].tolist()) == [10_000_000 + j for j in range(10)]
    p50, p95 = latency(idx, qp, 10, 200, 20)
    bytes_alg = idx.num_rows * 256.0
    emit({"case": "c4_shard", "pages": pages, "page_tokens": ptok, "store_gb": bytes_alg / 1e9, "fill_s": fill_s,
          "query": "1x16", "k": 10, "needles_found": needles_found, "gap_to_rank11": float(v[0, 9] - idx.search(qp, 11)[0][0, 10]),
          "p50_ms": p50, "p95_ms": p95, "hbm_floor_ms": bytes_alg / (HBM * 1e9) * 1e3,
          "frac_of_measured_hbm": bytes_alg / (p50 * 1e-3) / 1e9 / HBM, "gbs_at_p50": bytes_alg / (p50 * 1e-3) / 1e9})
    idx.close(); del idx; torch.cuda.empty_cache()

if "c3" in which:
    pages = 1_000_000
    lens = torch.randint(256, 769, (pages,), generator=torch.Generator().manual_seed(3003)).to(torch.int32).numpy()
    rows = int(lens.sum())
    idx = lis.LateInteractionIndex(rows, pages, device=dev)
    idx.fill_synthetic(pages, lens, seed=2003, id_base=0)
    torch.cuda.synchronize()
    bytes_alg = rows * 256.0
    for nq in (1, 32):
        qp = queries(nq, 32, 1003).pin_memory()
        p50, p95 = latency(idx, qp, 100, 30 if nq == 1 else 10, 5)
        flops = 2.0 * nq * 32 * 128 * rows
        emit({"case": "c3_full_1M_ragged", "pages": pages, "rows": rows, "store_gb": bytes_alg / 1e9, "nq": nq, "qtok": 32,
              "k": 100, "p50_ms": p50, "p95_ms": p95, "pairs_per_s": nq * pages / (p50 * 1e-3),
              "gbs": bytes_alg / (p50 * 1e-3) / 1e9, "frac_of_measured_hbm": bytes_alg / (p50 * 1e-3) / 1e9 / HBM,
              "tflops": flops / (p50 * 1e-3) / 1e12, "hbm_floor_ms": bytes_alg / (HBM * 1e9) * 1e3})
    idx.close(); del idx; torch.cuda.empty_cache()

if "c5" in which:
    pages, ptok, nq, qtok = 200_000, 1030, 1024, 32
    idx = lis.LateInteractionIndex(pages * ptok, pages, device=dev)
    idx.fill_synthetic(pages, ptok, seed=2005, id_base=0)
    store = idx._as_store()
    q = queries(nq, qtok, 1005).to(dev)
    pq = scoring.pack_queries(q, dev)
    scores = torch.empty((nq, pages), dtype=torch.float32, device=dev)
    best, mean = device_ms(lambda: scoring.maxsim_scores_device(pq, store, "f32", out=scores), iters=3, warm=1)
    flops = 2.0 * nq * qtok * 128 * pages * ptok
    emit({"case": "c5_full", "pages": pages, "nq": nq, "qtok": qtok, "ms_best": best, "ms_mean": mean,
          "pairs_per_s": nq * pages / (mean * 1e-3), "tflops": flops / (mean * 1e-3) / 1e12,
          "frac_of_measured_bf16_burst": flops / (mean * 1e-3) / 1e12 / TF, "score_matrix_gb": nq * pages * 4 / 1e9})
    # top-100 over the full matrix (K2 at scale)
    best, mean = device_ms(lambda: lis.topk_device(scores, 100), iters=3, warm=1)
    emit({"case": "c5_topk100", "nq": nq, "pages": pages, "ms_best": best, "ms_mean": mean})
    idx.close()
