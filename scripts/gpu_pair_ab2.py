"""Steady-state A/B of K1 forms on one GPU: modes rotate, every measurement = 20 untimed + 60 timed passes.
Usage: gpu_pair_ab2.py nq qtok [pages]"""
import importlib, json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
native = importlib.import_module("multi-modal_colpali_b200._native")
scoring = importlib.import_module("multi-modal_colpali_b200.scoring")
lib = native.load()
dev = torch.device("cuda", 0)
nq, qtok = int(sys.argv[1]), int(sys.argv[2])
pages = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
idx = lis.LateInteractionIndex(pages * 1030, pages, device=dev)
idx.fill_synthetic(pages, 1030, seed=7)
store = idx._as_store()
modes = [("single", (0, 0, 0, 0, 1)), ("pair", (0, 0, 0, 0, 3))]
q = torch.nn.functional.normalize(torch.randn(nq, qtok, 128, generator=torch.Generator().manual_seed(1)), dim=-1).to(torch.bfloat16).to(dev)
pq = scoring.pack_queries(q, dev)
scores = torch.empty((nq, pages), dtype=torch.float32, device=dev)
for rnd in range(4):
    order = modes if rnd % 2 == 0 else modes[::-1]
    for name, tun in order:
        native.check(lib.lis_set_tuning(*tun))
        for _ in range(20):
            scoring.maxsim_scores_device(pq, store, "f32", out=scores)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(60):
            scoring.maxsim_scores_device(pq, store, "f32", out=scores)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 60
        print(json.dumps({"round": rnd, "mode": name, "rows": nq * qtok, "pages": pages, "ms": round(ms, 4),
                          "tflops": round(2.0 * nq * qtok * 128 * pages * 1030 / ms / 1e9, 1)}), flush=True)
lib.lis_set_tuning(0, 0, 0, 0, 0)
