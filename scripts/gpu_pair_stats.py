"""Cycle attribution inside cluster 0 of the CTA-pair kernel (needs the LIS_K1_STATS build: LIS_LIB=.../liblis_stats.so)."""
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
pages = 30_000
idx = lis.LateInteractionIndex(pages * 1030, pages, device=dev)
idx.fill_synthetic(pages, 1030, seed=7)
store = idx._as_store()
stats = torch.zeros(256, dtype=torch.int64, device=dev)
for nq, qtok in [(8, 32), (12, 32), (32, 20), (24, 32)]:
    q = torch.nn.functional.normalize(torch.randn(nq, qtok, 128, generator=torch.Generator().manual_seed(1)), dim=-1).to(torch.bfloat16).to(dev)
    pq = scoring.pack_queries(q, dev)
    scores = torch.empty((nq, pages), dtype=torch.float32, device=dev)
    for name, tun in [("single", (0, 0, 0, 0, 1)), ("pair", (0, 0, 0, 0, 3))]:
        native.check(lib.lis_set_tuning(*tun))
        for _ in range(3):
            scoring.maxsim_scores_device(pq, store, "f32", out=scores)
        stats.zero_()
        native.check(lib.lis_k1_stats(stats.data_ptr()))
        scoring.maxsim_scores_device(pq, store, "f32", out=scores)
        torch.cuda.synchronize()
        native.check(lib.lis_k1_stats(None))
        s = stats.tolist()
        uses = max(s[3], 1)
        rec = {"rows": nq * qtok, "mode": name, "uses_counted": s[3], "mma_loop_per_use": round(s[0] / uses),
               "mma_wait_tiles_per_use": round(s[1] / uses), "mma_wait_acc_per_use": round(s[2] / uses),
               "mma_issue_per_use": round(s[23] / uses),
               "epi_wait_per_use": [round(s[4 + 2 * w] / uses) for w in range(8)],
               "epi_hold_per_use": [round(s[5 + 2 * w] / uses) for w in range(8)]}
        if name == "pair":
            rec.update({"peer_epi_wait_per_use": [round(s[64 + 4 + 2 * w] / uses) for w in range(8)],
                        "peer_epi_hold_per_use": [round(s[64 + 5 + 2 * w] / uses) for w in range(8)],
                        "producer_wait_total": [s[24], s[64 + 24]], "producer_loop_total": [s[25], s[64 + 25]],
                        "epi_loop_total": [s[26], s[64 + 26]], "mma_loop_total": s[0],
                        "split_use_hold_wait_n_w0": [s[28] / max(s[29], 1), s[30] / max(s[29], 1), s[29]],
                        "split_use_hold_wait_n_w4": [s[32] / max(s[33], 1), s[34] / max(s[33], 1), s[33]],
                        "all_use_hold_w0_w4_per_own_use": [2 * s[5] / uses, 2 * s[5 + 8] / uses],
                        "all_use_wait_w0_w4_per_own_use": [2 * s[4] / uses, 2 * s[4 + 8] / uses],
                        "slow_tiles_w0": {"cycles_per_slow_tile": s[40] / max(s[41], 1), "n_slow": s[41], "n_tiles": s[43],
                                          "fin_wait_total": s[42], "share_of_loop": s[40] / max(s[26], 1)},
                        "slow_tiles_w4": {"cycles_per_slow_tile": s[44] / max(s[45], 1), "n_slow": s[45], "n_tiles": s[47],
                                          "fin_wait_total": s[46], "share_of_loop": s[44] / max(s[26], 1)},
                        "spe_per_slow_tile_w0": {"take": s[48] / max(s[41], 1), "compute": s[49] / max(s[41], 1), "finish": s[50] / max(s[41], 1)},
                        "spe_per_slow_tile_w4": {"take": s[52] / max(s[45], 1), "compute": s[53] / max(s[45], 1), "finish": s[54] / max(s[45], 1)}})
        print(json.dumps(rec), flush=True)
lib.lis_set_tuning(0, 0, 0, 0, 0)
