"""Randomised stress of the search path (LateInteractionIndex: K1 + K2, incremental add, zero-padding blocks),
of fp32 embeddings (two bf16 planes) and of the projection head (K3) against the CPU oracle.
Usage: gpu_fuzz_search.py [cases] [seed]"""
import importlib, sys
from pathlib import Path
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
from oracle import maxsim_oracle as oracle   # checker (test driver, not product code)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
g = torch.Generator().manual_seed(int(sys.argv[2]) if len(sys.argv) > 2 else 7)


def unit(x):
    return x / x.norm(dim=-1, keepdim=True).clamp_min(1e-20)


def rint(lo, hi):
    return int(torch.randint(lo, hi + 1, (1,), generator=g))


bad = 0
for c in range(cases):
    kind = ["search", "search", "fp32", "k3"][rint(0, 3)]
    ok, note = True, ""
    if kind == "search":
        dtype = torch.bfloat16 if rint(0, 2) else torch.float16
        nq = rint(1, 48)
        qs = [unit(torch.randn(rint(1, 70), 128, generator=g)).to(dtype) for _ in range(nq)]
        n_pages = rint(1, 1500)
        p_lens = [max(1, rint(-200, 900)) for _ in range(n_pages)]
        ps = [unit(torch.randn(n, 128, generator=g)).to(dtype) for n in p_lens]
        k = rint(1, min(150, n_pages + 20))
        blk = [None, 128, 16][rint(0, 2)]
        ids = torch.randperm(10 * n_pages, generator=g)[:n_pages].tolist()
        idx = lis.LateInteractionIndex(sum(p_lens) + 8, n_pages + 1, dtype=dtype)
        cut = rint(0, n_pages)                                   # two incremental adds
        if blk is None:
            if cut: idx.add(ps[:cut], ids=ids[:cut])
            if cut < n_pages: idx.add(ps[cut:], ids=ids[cut:])
        else:
            idx.add(ps, ids=ids, zero_pad_block=blk)
        v, i = idx.search(qs, k)
        want = oracle.score_multi_vector_widened(qs, ps, batch_size=blk or 10**9) if blk else oracle.score_multi_vector_widened(
            qs, [p for p in ps], batch_size=1)                    # no zero padding semantics: every page its own block
        id_t = torch.tensor(ids)
        kk = min(k, n_pages)
        for q in range(nq):
            row = want[q]
            best = torch.sort(row, descending=True).values[:kk]
            got_ids = i[q, :kk].tolist()
            pos = {pid: j for j, pid in enumerate(ids)}
            if len(set(got_ids)) != kk or any(pid not in pos for pid in got_ids):
                ok = False; note = f"bad ids for query {q}"; break
            got_sc = torch.tensor([row[pos[pid]].item() for pid in got_ids])
            if (got_sc - v[q, :kk]).abs().max() > 1e-4 or (best - v[q, :kk]).abs().max() > 1e-4:
                ok = False; note = f"scores off for query {q}: {(best - v[q, :kk]).abs().max().item():.2e}"; break
            if k > n_pages and not ((i[q, kk:] == -1).all() and torch.isinf(v[q, kk:]).all()):
                ok = False; note = "padding tail wrong"; break
        idx.close()
        note = note or f"nq={nq} pages={n_pages} k={k} block={blk} {str(dtype)[6:]}"
    elif kind == "fp32":
        nq = rint(1, 12)
        qs = [unit(torch.randn(rint(1, 200), 128, generator=g)) for _ in range(nq)]
        ps = [unit(torch.randn(max(1, rint(-50, 700)), 128, generator=g)) for _ in range(rint(1, 300))]
        bs = [128, 16][rint(0, 1)]
        got = lis.score_multi_vector(qs, ps, batch_size=bs)
        want = oracle.score_multi_vector(qs, ps, batch_size=bs)
        err = (got - want).abs().max().item()
        ok = err <= 1e-4
        note = f"nq={nq} pages={len(ps)} err={err:.2e}"
    else:
        n_tok, hidden = rint(1, 5000), 64 * rint(1, 40)
        h = torch.randn(n_tok, hidden, generator=g).to(torch.bfloat16)
        w = (torch.randn(128, hidden, generator=g) / hidden ** 0.5).to(torch.bfloat16)
        b = (0.1 * torch.randn(128, generator=g)).to(torch.bfloat16) if rint(0, 1) else None
        m = (torch.rand(n_tok, generator=g) > 0.2).to([torch.int64, torch.int32, torch.uint8][rint(0, 2)])
        got = lis.project_normalize(h.cuda(), w.cuda(), None if b is None else b.cuda(), m.cuda()).float().cpu()
        want = oracle.project_normalize(h.float(), w.float(), None if b is None else b.float(), m).float()
        err = (got - want).abs().max().item()
        ok = err <= 4e-3                                           # one bf16 step of a unit-norm component (|x| < 1)
        note = f"n_tok={n_tok} hidden={hidden} err={err:.2e}"
    bad += 0 if ok else 1
    print(f"case {c} [{kind}] {note} {'ok' if ok else 'MISMATCH'}", flush=True)
print("FUZZ", "PASS" if bad == 0 else f"FAIL ({bad})")
sys.exit(0 if bad == 0 else 1)
