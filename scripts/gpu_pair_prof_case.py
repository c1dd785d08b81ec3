"""Small K1 case for ncu: 20 000 pages x 1030, 5 and 6 query tiles (default tuning)."""
import importlib, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
scoring = importlib.import_module("multi-modal_colpali_b200.scoring")
dev = torch.device("cuda", 0)
pages = 20_000
idx = lis.LateInteractionIndex(pages * 1030, pages, device=dev)
idx.fill_synthetic(pages, 1030, seed=7)
store = idx._as_store()
for nq, qtok in [(32, 20), (24, 32)]:
    q = torch.nn.functional.normalize(torch.randn(nq, qtok, 128, generator=torch.Generator().manual_seed(1)), dim=-1).to(torch.bfloat16).to(dev)
    pq = scoring.pack_queries(q, dev)
    scores = torch.empty((nq, pages), dtype=torch.float32, device=dev)
    for _ in range(3):
        scoring.maxsim_scores_device(pq, store, "f32", out=scores)
    torch.cuda.synchronize()
print("ok")
