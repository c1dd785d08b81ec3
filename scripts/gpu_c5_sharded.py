"""BASELINE configs[4] on N GPUs (torchrun): 1024 queries x 32 tokens vs 200 000 pages x 1030 tokens, corpus sharded by
page; every rank computes its column block of the full [1024, 200k] score matrix (no data-path collective).
Time = max over ranks of the CUDA-event duration.  Rank 0 appends a JSON line to gpurun_out/c5_sharded.jsonl."""
import importlib, json, os, sys
from pathlib import Path
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
scoring = importlib.import_module("multi-modal_colpali_b200.scoring")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}

pages, ptok, nq, qtok = 200_000, 1030, 1024, 32
a, b = lis.shard_range(pages, rank, world)
idx = lis.LateInteractionIndex((b - a) * ptok, b - a, device=dev)
idx.fill_synthetic(b - a, ptok, seed=2005 + rank, id_base=a)
store = idx._as_store()
q = torch.nn.functional.normalize(torch.randn(nq, qtok, 128, generator=torch.Generator().manual_seed(1005)), dim=-1).to(torch.bfloat16).to(dev)
pq = scoring.pack_queries(q, dev)
scores = torch.empty((nq, b - a), dtype=torch.float32, device=dev)
for _ in range(2):
    scoring.maxsim_scores_device(pq, store, "f32", out=scores)
ts = []
for _ in range(4):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); scoring.maxsim_scores_device(pq, store, "f32", out=scores); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = torch.tensor(ts, dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.mean().item())
    flops = 2.0 * nq * qtok * 128 * pages * ptok
    rec = {"case": "c5_sharded", "n_gpus": world, "pages": pages, "nq": nq, "qtok": qtok, "ms_mean_max_over_ranks": ms,
           "pairs_per_s": nq * pages / (ms * 1e-3), "tflops_total": flops / (ms * 1e-3) / 1e12,
           "tflops_per_gpu": flops / (ms * 1e-3) / 1e12 / world,
           "frac_of_measured_bf16_burst": flops / (ms * 1e-3) / 1e12 / world / pk["bf16_tflops"],
           "frac_of_measured_bf16_sustained": flops / (ms * 1e-3) / 1e12 / world / pk["bf16_tflops_sustained"]}
    print(json.dumps(rec), flush=True)
    with open(ROOT / "gpurun_out" / "c5_sharded.jsonl", "a") as f:
        f.write(json.dumps(rec) + "\n")
dist.destroy_process_group()
