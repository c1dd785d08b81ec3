"""Where does K1's time go?  Same launch with parts of the epilogue switched off (results invalid)."""
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
pages = 50_000
idx = lis.LateInteractionIndex(pages * 1030, pages, device=dev)
idx.fill_synthetic(pages, 1030, seed=7)
store = idx._as_store()
q = torch.nn.functional.normalize(torch.randn(12, 32, 128, generator=torch.Generator().manual_seed(1)), dim=-1).to(torch.bfloat16).to(dev)
pq = scoring.pack_queries(q, dev)      # 384 rows = 3 M tiles, one pass
scores = torch.empty((12, pages), dtype=torch.float32, device=dev)
flops = 2.0 * 384 * 128 * pages * 1030


def t(iters=6):
    for _ in range(2):
        scoring.maxsim_scores_device(pq, store, "f32", out=scores)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); scoring.maxsim_scores_device(pq, store, "f32", out=scores); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


for tiling in [(256, 3, 2, 1), (128, 3, 2, 2), (128, 3, 2, 1)]:
    for mode, name in [(0, "full"), (2, "no max"), (1, "no tmem loads"), (0, "full again")]:
        native.check(lib.lis_set_tuning(*tiling[:2], 0, tiling[2], tiling[3]))
        native.check(lib.lis_set_ablation(mode))
        best, mean = t()
        print(json.dumps({"tiling": tiling, "mode": name, "ms_best": best, "ms_mean": mean, "tflops_best": flops / best / 1e9}), flush=True)
lib.lis_set_ablation(0)
