"""A/B on one GPU: single-CTA K1 (a_operand=1) vs the CTA-pair form (a_operand=3), alternating so drift cancels.
Usage: gpu_pair_ab.py [pages]"""
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
pages = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000
idx = lis.LateInteractionIndex(pages * 1030, pages, device=dev)
idx.fill_synthetic(pages, 1030, seed=7)
store = idx._as_store()
out = open(ROOT / "gpurun_out" / "pair_ab.jsonl", "a")
cases = [(8, 32), (12, 32), (32, 20), (24, 32), (64, 32)]
modes = [("single", (0, 0, 0, 0, 1)), ("pair6", (0, 0, 0, 0, 3)), ("pair7", (0, 7, 0, 0, 3)), ("pair4", (0, 4, 0, 0, 3))]
for rnd in range(2):
    for nq, qtok in cases:
        q = torch.nn.functional.normalize(torch.randn(nq, qtok, 128, generator=torch.Generator().manual_seed(1)), dim=-1).to(torch.bfloat16).to(dev)
        pq = scoring.pack_queries(q, dev)
        scores = torch.empty((nq, pages), dtype=torch.float32, device=dev)
        for name, tun in modes:
            tiles = (nq * qtok + 127) // 128
            if name == "pair7" and tiles < 7: continue
            if name == "pair4" and tiles <= 4: continue
            native.check(lib.lis_set_tuning(*tun))
            for _ in range(3):
                scoring.maxsim_scores_device(pq, store, "f32", out=scores)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                scoring.maxsim_scores_device(pq, store, "f32", out=scores)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            rec = {"round": rnd, "mode": name, "rows": nq * qtok, "tiles": tiles, "pages": pages, "ms": round(ms, 4),
                   "tflops": round(2.0 * nq * qtok * 128 * pages * 1030 / ms / 1e9, 1),
                   "gbs_one_pass": round(pages * 1030 * 256 / ms / 1e6, 1)}
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n"); out.flush()
lib.lis_set_tuning(0, 0, 0, 0, 0)
