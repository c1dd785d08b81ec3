"""Multi-GPU invariance check (run under torchrun, one rank per GPU; `tests/test_gpu_round2.py::test_multi_gpu_world_invariance`
launches it for every world size the box offers).  Each rank holds a contiguous page shard (balanced by token count) with
global page ids and ONE call -- `ShardedIndex.search` -> `lis_index_search_sharded` (K1, K2, ncclAllGather, merge, download,
replayed as a CUDA graph) -- must return, on every rank and for any world size, exactly the top-k of the unsharded corpus:
bit-identical to a single-GPU search of the whole corpus and to the CPU oracle's (score desc, id asc) order.

Prints `RESULT {...}` and `sharded-check PASS` on rank 0; exits non-zero on any mismatch.
"""
import importlib
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lis = importlib.import_module("multi-modal_colpali_b200")
    from oracle import maxsim_oracle as oracle

    g = torch.Generator().manual_seed(11)
    unit = lambda x: x / x.norm(dim=-1, keepdim=True)
    n_pages = 4001
    lens = torch.randint(8, 96, (n_pages,), generator=g).tolist()
    lens[7] = 0                                   # an empty page
    pages = [unit(torch.randn(n, 128, generator=g)).to(torch.bfloat16) for n in lens]
    qs = [unit(torch.randn(n, 128, generator=g)).to(torch.bfloat16) for n in (16, 20, 32, 100, 7)]
    pages[1234] = torch.cat([qs[0], pages[1234]])[: max(lens[1234], 16)]     # a planted needle
    pages[4000] = pages[17].clone()               # an exact tie across shards: (score desc, id asc) must decide
    lens = [int(p.shape[0]) for p in pages]
    full = oracle.score_multi_vector_widened(qs, pages, batch_size=10 ** 9)
    ok = True
    report = {"world": world, "cases": []}
    for name, ranges in (("balanced", lis.balanced_shard_ranges(lens, world)),
                         ("rank0-empty", [(0, 0)] + [lis.shard_range(n_pages, r, world - 1) for r in range(world - 1)]
                          if world > 1 else [(0, n_pages)])):
        a, b = ranges[rank]
        idx = lis.LateInteractionIndex(max(sum(lens[a:b]), 1), max(b - a, 1), device=dev)
        if b > a:
            # zero-padding semantics like the oracle's single block: shorter pages (the empty one included) clamp at 0
            idx.add(pages[a:b], ids=list(range(a, b)), zero_pad_block=10 ** 9)
        sh = lis.ShardedIndex(idx)
        for k in (10, 100):
            want_v, want_i = oracle.topk(full, k)
            for rep in range(3):                  # first call runs eagerly and captures; the next ones replay the graph
                v, i = sh.search(qs, k)
                same = torch.equal(i, want_i) and (v - want_v).abs().max().item() <= 1e-4
                ok = ok and same
            # every rank must hold the same merged answer
            mine = torch.cat([v.flatten().view(torch.int32).to(torch.int64), i.flatten()]).to(dev)
            ref = mine.clone()
            dist.broadcast(ref, src=0)
            ok = ok and bool(torch.equal(mine, ref))
            report["cases"].append({"sharding": name, "k": k, "ok": bool(same), "graphs": idx.graph_stats()})
        # the torch-level path (device tensors, gather_candidates) must agree with the C path
        v2, i2 = sh.search_device(qs, 10)
        v1, i1 = sh.search(qs, 10)
        ok = ok and torch.equal(i2.cpu(), i1) and torch.equal(v2.cpu(), v1)
        sh.close()
        idx.close()
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        report["ok"] = int(flag.item()) == 0
        print("RESULT " + json.dumps(report), flush=True)
        print("sharded-check PASS" if report["ok"] else "sharded-check FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
