import importlib, json, os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
scoring = importlib.import_module("multi-modal_colpali_b200.scoring")
dev = torch.device("cuda", 0)
pages = 60_000
idx = lis.LateInteractionIndex(pages * 1030, pages, device=dev)
idx.fill_synthetic(pages, 1030, seed=7)
store = idx._as_store()
res = {}
for nq, qtok in ((8, 32), (32, 20), (24, 32), (64, 32)):
    q = torch.nn.functional.normalize(torch.randn(nq, qtok, 128, generator=torch.Generator().manual_seed(1)), dim=-1).to(torch.bfloat16).to(dev)
    pq = scoring.pack_queries(q, dev)
    scores = torch.empty((nq, pages), dtype=torch.float32, device=dev)
    for _ in range(3):
        scoring.maxsim_scores_device(pq, store, "f32", out=scores)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        scoring.maxsim_scores_device(pq, store, "f32", out=scores)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    res[f"{nq * qtok}rows"] = round(2.0 * nq * qtok * 128 * pages * 1030 / ms / 1e9, 1)
print(json.dumps({"lib": os.environ.get("LIS_LIB", "default"), "tflops": res}), flush=True)
