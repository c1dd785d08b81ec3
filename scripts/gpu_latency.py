"""Search latency / throughput at the corpus sizes the reference actually uses (hundreds of pages) and
the effect of coalescing concurrent single-query searches (QueryBatcher) on a large shard."""
import importlib, json, sys, threading, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
dev = torch.device("cuda", 0)
out = open(ROOT / "gpurun_out" / "latency.jsonl", "a")


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def emit(rec):
    print(json.dumps(rec), flush=True); out.write(json.dumps(rec) + "\n"); out.flush()


g = torch.Generator().manual_seed(3)
for pages in (300, 3000, 30000):
    idx = lis.LateInteractionIndex(pages * 1030, pages, device=dev)
    idx.fill_synthetic(pages, 1030, seed=11)
    for nq in (1, 10):
        q = unit(torch.randn(nq, 20, 128, generator=g)).to(torch.bfloat16).pin_memory()
        for _ in range(20):
            idx.search(q, 5)
        ts = []
        for _ in range(200):
            t0 = time.perf_counter(); idx.search(q, 5); ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        emit({"case": "small_corpus_search", "pages": pages, "nq": nq, "k": 5, "p50_ms": ts[100], "p95_ms": ts[190],
              "store_mb": pages * 1030 * 256 / 1e6})
    idx.close()

# coalescing on a big shard: 200k pages x 1030 (52.7 GB): 16 client threads, each 20 single-query searches
pages = 200_000
idx = lis.LateInteractionIndex(pages * 1030, pages, device=dev)
idx.fill_synthetic(pages, 1030, seed=12)
qs = [unit(torch.randn(16, 128, generator=g)).to(torch.bfloat16) for _ in range(16)]
for _ in range(3):
    idx.search([qs[0]], 10)
t0 = time.perf_counter()
for i in range(32):
    idx.search([qs[i % 16]], 10)
seq = (time.perf_counter() - t0) / 32
emit({"case": "sequential_single_query", "pages": pages, "ms_per_query": seq * 1e3, "qps": 1 / seq})
# what one coalesced pass costs as it grows (no threads: index.search on a list of n single queries of 16 tokens):
# up to 3 query tiles run on one CTA per SM, 4+ tiles on CTA pairs
qs80 = [unit(torch.randn(16, 128, generator=g)).to(torch.bfloat16) for _ in range(80)]
for n in (1, 8, 16, 24, 32, 48, 64, 80):
    for _ in range(3):
        idx.search(qs80[:n], 10)
    ts = []
    for _ in range(12):
        t0 = time.perf_counter(); idx.search(qs80[:n], 10); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    emit({"case": "one_pass_n_queries", "pages": pages, "queries": n, "query_rows": 16 * n, "p50_ms": ts[6], "qps": n / (ts[6] * 1e-3)})
for max_rows in (128, 256):
    b = lis.QueryBatcher(idx, max_rows=max_rows, max_wait_ms=0.5)
    lat = []
    lock = threading.Lock()

    def client(j):
        for _ in range(20):
            t = time.perf_counter(); b.search(qs[j], 10); d = time.perf_counter() - t
            with lock:
                lat.append(d * 1e3)

    th = [threading.Thread(target=client, args=(j,)) for j in range(16)]
    t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]
    wall = time.perf_counter() - t0
    lat.sort()
    emit({"case": "coalesced_16_clients", "pages": pages, "max_rows": max_rows, "qps": 320 / wall, "p50_ms": lat[len(lat) // 2],
          "p95_ms": lat[int(len(lat) * 0.95)], "passes": b.batches, "queries": b.served})
    b.close()
