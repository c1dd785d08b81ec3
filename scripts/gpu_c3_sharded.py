"""BASELINE configs[2] on N GPUs (torchrun): 1 000 000 ragged pages (256..768 tokens, ~131 GB in total) sharded by
token count, 1 x 32 and 32 x 32 query tokens, top-100 with one NCCL all-gather + merge.  Rank 0 appends JSON
lines to gpurun_out/c3_sharded.jsonl; latency = max over ranks of the host-side wall time per search."""
import importlib, json, os, sys, time
from pathlib import Path
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
HBM = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0

pages = 1_000_000
lens = torch.randint(256, 769, (pages,), generator=torch.Generator().manual_seed(3003)).to(torch.int32).numpy()
a, b = lis.balanced_shard_ranges(lens, world)[rank]
my = lens[a:b]
rows = int(my.sum())
idx = lis.LateInteractionIndex(rows, b - a, device=dev)
idx.fill_synthetic(b - a, my, seed=2003 + rank, id_base=a)
torch.cuda.synchronize()
sharded = lis.ShardedIndex(idx)
max_rows = torch.tensor([rows], dtype=torch.int64, device=dev)
dist.all_reduce(max_rows, op=dist.ReduceOp.MAX)


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


for nq, iters in ((1, 100), (32, 30)):
    q = unit(torch.randn(nq, 32, 128, generator=torch.Generator().manual_seed(1003))).to(torch.bfloat16).pin_memory()
    for _ in range(5):
        sharded.search(q, 100)
    ts = []
    for _ in range(iters):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        v, i = sharded.search(q, 100)
        ts.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor(ts, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts = sorted(t.tolist())
    if rank == 0:
        p50 = ts[len(ts) // 2]
        floor = int(max_rows.item()) * 256.0 / (HBM * 1e9) * 1e3
        rec = {"case": "c3_sharded_1M_ragged", "n_gpus": world, "pages": pages, "rows_largest_shard": int(max_rows.item()),
               "nq": nq, "qtok": 32, "k": 100, "p50_ms": p50, "p95_ms": ts[int(len(ts) * 0.95)], "pairs_per_s": nq * pages / (p50 * 1e-3),
               "hbm_floor_ms_largest_shard": floor, "tflops_total": 2.0 * nq * 32 * 128 * float(lens.sum()) / (p50 * 1e-3) / 1e12,
               "ids_unique": len(set(i[0].tolist())) == 100}
        print(json.dumps(rec), flush=True)
        with open(ROOT / "gpurun_out" / "c3_sharded.jsonl", "a") as f:
            f.write(json.dumps(rec) + "\n")
dist.destroy_process_group()
