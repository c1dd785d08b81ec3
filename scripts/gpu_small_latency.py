"""Host-to-host latency of the one-shot search at the reference's own corpus size (300 pages), by query container:
a [10, 20, 128] tensor, a list of 10 tensors, one query.  A few seconds on one GPU."""
import importlib, json, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
dev = torch.device("cuda", 0)
idx = lis.LateInteractionIndex(300 * 1030, 300, device=dev)
idx.fill_synthetic(300, 1030, seed=7, id_base=0)
q = torch.nn.functional.normalize(torch.randn(10, 20, 128, generator=torch.Generator().manual_seed(3)), dim=-1).to(torch.bfloat16)
ql = [q[i].clone() for i in range(10)]
want = idx.search(q, 5)
for name, arg in (("tensor_10x20", q), ("list_of_10", ql), ("one_query", [ql[0]])):
    got = idx.search(arg, 5)
    n = got[0].shape[0]
    assert torch.equal(got[0], want[0][:n]) and torch.equal(got[1], want[1][:n]), name
    lat = []
    for i in range(520):
        t0 = time.perf_counter()
        idx.search(arg, 5)
        lat.append((time.perf_counter() - t0) * 1e6)
    lat = sorted(lat[20:])
    print(json.dumps({"queries": name, "pages": 300, "k": 5, "p50_us": round(lat[len(lat) // 2], 1),
                      "p05_us": round(lat[len(lat) // 20], 1), "p95_us": round(lat[len(lat) * 19 // 20], 1),
                      "graphs": idx.graph_stats()}), flush=True)
idx.close()
