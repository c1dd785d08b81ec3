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
for rnd in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    for mode, name in [(0, "full"), (2, "no max"), (1, "no tmem loads, no max"), (3, "no TMA traffic"),
                       (4, "no TMA traffic, no tmem loads, no max")]:
        native.check(lib.lis_set_ablation(mode))
        ms = run()
        print(json.dumps({"round": rnd, "mode": name, "ms": ms, "tflops": flops / ms / 1e9}), flush=True)
lib.lis_set_ablation(0)
# cycle attribution inside CTA 0 (one launch per mode); needs the LIS_K1_STATS build (LIS_LIB=.../liblis_stats.so)
import os
if "stats" not in os.environ.get("LIS_LIB", ""):
    sys.exit(0)
stats = torch.zeros(256, dtype=torch.int64, device=dev)
for mode, name in [(0, "full"), (1, "no tmem loads, no max"), (3, "no TMA traffic"), (4, "no TMA traffic, no tmem loads, no max")]:
    native.check(lib.lis_set_ablation(mode))
    run(5)
    stats.zero_()
    native.check(lib.lis_k1_stats(stats.data_ptr()))
    scoring.maxsim_scores_device(pq, store, "f32", out=scores)
    torch.cuda.synchronize()
    native.check(lib.lis_k1_stats(None))
    s = stats.tolist()
    uses = max(s[3], 1)
    t0 = min(x for x in s[32:72] if x > 0) if any(s[32:72]) else 0
    tl = [[(s[32 + u * 5 + k] - t0) if s[32 + u * 5 + k] else None for k in range(5)] for u in range(8)]
    print("timeline (cycles; per use: mma_wait_start, mma_wait_end, mma_issued, epi_wake, epi_release):", name)
    for u, row in enumerate(tl):
        wk = [s[96 + u * 16 + w] - t0 for w in range(8)]
        rl = [s[96 + u * 16 + 8 + w] - t0 for w in range(8)]
        print("   use", 1000 + u, row, "wake by warp", wk, "release by warp", rl)
    print(json.dumps({"mode": name, "mma_loop_cycles_per_use": s[0] / uses, "mma_wait_tiles_per_use": s[1] / uses,
                      "mma_wait_acc_per_use": s[2] / uses, "mma_issue_per_use": s[23] / uses,
                      "epi_wait_full_per_use_by_warp": [round(s[4 + 2 * w] / uses) for w in range(8)],
                      "epi_hold_per_use_by_warp": [round(s[5 + 2 * w] / uses) for w in range(8)], "uses": s[3],
                      "w0_finish_cycles_each": s[20] / max(s[21], 1), "w0_finishes": s[21], "w0_post_release_per_use": s[22] / uses}), flush=True)
lib.lis_set_ablation(0)
