"""Steady-state cost of ONE pass over the page store for every K1 form / resident tile count
(the table behind the pass planner in lis_maxsim.cu).  60 000 pages x 1030 tokens."""
import importlib, json, os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
native = importlib.import_module("multi-modal_colpali_b200._native")
scoring = importlib.import_module("multi-modal_colpali_b200.scoring")
lib = native.load()
dev = torch.device("cuda", 0)
pages = 60_000
idx = lis.LateInteractionIndex(pages * 1030, pages, device=dev)
idx.fill_synthetic(pages, 1030, seed=7)
store = idx._as_store()
cases = [("single", n, (0, n, 0, 0, 1)) for n in (1, 2, 3)] + [("pair", n, (0, n, 0, 0, 3)) for n in (2, 3, 4, 5, 6, 7, 8, 9, 10)]
for rnd in range(int(os.environ.get('ROUNDS', '2'))):
    for name, n, tun in (cases if rnd == 0 else cases[::-1]):
        q = torch.nn.functional.normalize(torch.randn(n * 4, 32, 128, generator=torch.Generator().manual_seed(1)), dim=-1).to(torch.bfloat16).to(dev)
        pq = scoring.pack_queries(q, dev)
        scores = torch.empty((n * 4, pages), dtype=torch.float32, device=dev)
        native.check(lib.lis_set_tuning(*tun))
        for _ in range(25):
            scoring.maxsim_scores_device(pq, store, "f32", out=scores)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            scoring.maxsim_scores_device(pq, store, "f32", out=scores)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        print(json.dumps({"round": rnd, "form": name, "tiles": n, "ms": round(ms, 4), "ms_per_tile": round(ms / n, 4),
                          "tflops": round(2.0 * n * 128 * 128 * pages * 1030 / ms / 1e9, 1),
                          "gbs": round(pages * 1030 * 256 / ms / 1e6, 1)}), flush=True)
lib.lis_set_tuning(0, 0, 0, 0, 0)
