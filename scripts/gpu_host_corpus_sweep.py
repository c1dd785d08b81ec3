"""Host-resident corpus route (lis_stream_scores): chunk size x gather threads, pageable source, against the PCIe floor."""
import importlib, json, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
scoring = importlib.import_module("multi-modal_colpali_b200.scoring")
native = importlib.import_module("multi-modal_colpali_b200._native")
dev = torch.device("cuda", 0)
pages = 10_000
idx = lis.LateInteractionIndex(pages * 1030, pages, device=dev)
idx.fill_synthetic(pages, 1030, seed=5)
host = idx._as_store().tokens.view(pages, 1030, 128).cpu()
idx.close()
q = torch.nn.functional.normalize(torch.randn(32, 20, 128, generator=torch.Generator().manual_seed(1)), dim=-1).to(torch.bfloat16)
pq = scoring.pack_queries(q, dev)
pin = torch.empty(1 << 29, dtype=torch.uint8).pin_memory(); pin.zero_()
dst = torch.empty(1 << 29, dtype=torch.uint8, device=dev)
best = 1e9
for it in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dst.copy_(pin, non_blocking=True); e1.record(); torch.cuda.synchronize()
    if it >= 2: best = min(best, e0.elapsed_time(e1))
gbs = (1 << 29) / (best * 1e-3) / 1e9
floor = host.numel() * 2 / (gbs * 1e9) * 1e3
print(json.dumps({"h2d_gbs": gbs, "floor_ms": floor}), flush=True)
for chunk in (1 << 16, 1 << 17, 1 << 18, 1 << 19):
    for thr in (4, 8, 12, 16):
        ts = []
        for it in range(5):
            t0 = time.perf_counter()
            scoring.stream_scores_host_corpus(pq, host, 128, "f32", chunk_rows=chunk, host_threads=thr)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        ms = sorted(ts[1:])[len(ts[1:]) // 2]
        print(json.dumps({"chunk_rows": chunk, "threads": thr, "ms": round(ms, 2), "frac_of_floor": round(floor / ms, 3)}), flush=True)
native.load().lis_stream_release()
