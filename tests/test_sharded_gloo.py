"""world_size=2 over gloo on CPU: the shard -> local top-k -> all-gather -> merge plumbing gives the
same answer as searching the unsharded corpus.  The CUDA kernels are replaced by injected oracle
callables here (this is the host logic; the kernels themselves are covered by the -m gpu tests)."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, ret):
    sys.path.insert(0, str(ROOT))
    import importlib

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lis = importlib.import_module("multi-modal_colpali_b200")
    from oracle import maxsim_oracle as oracle

    g = torch.Generator().manual_seed(5)
    n_pages, k = 101, 7
    ps = [torch.randn(int(n), 128, generator=g) for n in torch.randint(3, 20, (n_pages,), generator=g)]
    qs = [torch.randn(6, 128, generator=g), torch.randn(11, 128, generator=g)]
    a, b = lis.shard_range(n_pages, rank, world)

    def local_search(qs_, k_, round_mode):
        s = oracle.score_multi_vector(qs_, ps[a:b], batch_size=10 ** 9)
        v, i = oracle.topk(s, k_)
        pad = k_ - v.shape[1]
        if pad > 0:
            v = torch.cat([v, torch.full((v.shape[0], pad), float("-inf"))], 1)
            i = torch.cat([i, torch.full((i.shape[0], pad), -1 - a, dtype=torch.int64)], 1)
        return v, i + a          # global page ids

    sharded = lis.ShardedIndex(None, local_search=local_search,
                               merge=lambda s, i, k_: oracle.merge_topk([(s, i)], k_))
    v, i = sharded.search(qs, k)
    full = oracle.score_multi_vector(qs, ps, batch_size=10 ** 9)
    wv, wi = oracle.topk(full, k)
    ok = torch.equal(i, wi) and torch.allclose(v, wv)
    # raw gather layout: rank-major columns
    s_loc = torch.full((2, 3), float(rank)); i_loc = torch.arange(3).repeat(2, 1) + 100 * rank
    gs, gi = lis.gather_candidates(s_loc, i_loc)
    ok = ok and gs.shape == (2, 3 * world) and gs[0].tolist() == [0.0] * 3 + [1.0] * 3 \
        and gi[1].tolist() == [0, 1, 2, 100, 101, 102]
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_sharded_search_equals_unsharded():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
