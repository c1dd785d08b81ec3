"""torchrun --nproc-per-node N scripts/sharded_check.py : the same synthetic corpus searched with world=N
shards must return exactly what one GPU returns for the whole corpus (scores and ids), top-10 and top-100."""
import importlib
import os
import sys
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

n_pages = 40_000
g = torch.Generator().manual_seed(3003)
lens = torch.randint(256, 769, (n_pages,), generator=g).tolist()
qs = [torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=-1).to(torch.bfloat16) for n in (32, 16, 20)]

# every rank can regenerate any page: rows are a pure function of (seed, global row), so build the
# full corpus on rank 0's GPU for the unsharded answer and only the local slice elsewhere
parts = lis.balanced_shard_ranges(lens, world)
a, b = parts[rank]
row_start = sum(lens[:a])
native = importlib.import_module("multi-modal_colpali_b200._native")
lib = native.load()


def build(lo, hi, first_row):
    idx = lis.LateInteractionIndex(sum(lens[lo:hi]), hi - lo, device=dev)
    # fill_synthetic numbers rows from the index's own row 0; give the hash the global row instead
    store_rows = sum(lens[lo:hi])
    idx.fill_synthetic(hi - lo, lens[lo:hi], seed=77, id_base=lo)
    tok = idx._as_store().tokens
    native.check(lib.lis_fill_synthetic_rows(tok.data_ptr(), first_row, store_rows, 77, 0,
                                             torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return idx


local_idx = build(a, b, row_start)
sharded = lis.ShardedIndex(local_idx)
ok = True
for k in (10, 100):
    v, i = sharded.search(qs, k)
    if rank == 0:
        full = build(0, n_pages, 0)
        wv, wi = full.search(qs, k)
        same = torch.equal(i, wi) and torch.equal(v, wv)
        print(f"world={world} k={k}: sharded == unsharded: {same}", flush=True)
        ok = ok and same
        full.close()
dist.barrier()
if rank == 0:
    print("SHARDED_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
