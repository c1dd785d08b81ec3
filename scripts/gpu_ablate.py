"""Where does K1's time (energy) go?  The same launch with parts of the epilogue switched off (results
invalid), alternating modes at steady state so that power-cap drift cancels."""
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
pages = 60_000
idx = lis.LateInteractionIndex(pages * 1030, pages, device=dev)
idx.fill_synthetic(pages, 1030, seed=7)
store = idx._as_store()
q = torch.nn.functional.normalize(torch.randn(12, 32, 128, generator=torch.Generator().manual_seed(1)), dim=-1).to(torch.bfloat16).to(dev)
pq = scoring.pack_queries(q, dev)      # 384 rows = 3 M tiles, one pass
scores = torch.empty((12, pages), dtype=torch.float32, device=dev)
flops = 2.0 * 384 * 128 * pages * 1030


def run(n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        scoring.maxsim_scores_device(pq, store, "f32", out=scores)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


run(40)  # reach the power-capped steady state
for rnd in range(3):
    for mode, name in [(0, "full"), (2, "no max"), (1, "no tmem loads, no max")]:
        native.check(lib.lis_set_ablation(mode))
        ms = run()
        print(json.dumps({"round": rnd, "mode": name, "ms": ms, "tflops": flops / ms / 1e9}), flush=True)
lib.lis_set_ablation(0)
