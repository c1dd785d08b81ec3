"""Small, fast invocation of every kernel (K1 single-CTA / CTA-pair incl. the page-end paths and the split tile, K2, K3 in
both rounding modes, fused ingestion, one-shot graph search, host-corpus streaming) for compute-sanitizer:

    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_case.py

Checks results against the oracle as well, so a sanitizer run is also a parity run."""
import importlib
import math
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

lis = importlib.import_module("multi-modal_colpali_b200")
from oracle import maxsim_oracle as oracle

g = torch.Generator().manual_seed(5)
unit = lambda x: x / x.norm(dim=-1, keepdim=True)
ru = lambda *s: unit(torch.randn(*s, generator=g)).to(torch.bfloat16)
worst = 0.0
# pages: ColPali-like (one page end per 256-row tile at most), short ragged ones (several ends per tile), empty ones
pages = [ru(n, 128) for n in [1030, 1030, 700, 1030, 3, 0, 17, 300, 257, 256, 1, 90, 0, 0, 511] * 3]
for nq, qt in [(1, 16), (10, 20), (32, 20), (24, 32), (28, 32), (36, 32)]:      # 1, 2, 5, 6, 7, 9 query tiles
    q = ru(nq, qt, 128)
    want = oracle.score_multi_vector_widened(q, pages)
    got = lis.score_multi_vector(q, [p.cuda() for p in pages], round_mode="f32")
    worst = max(worst, (got - want).abs().max().item())
    got_h = lis.score_multi_vector(q, pages, round_mode="f32")                   # host-resident corpus, streamed
    assert torch.equal(got, got_h)
assert worst <= 1e-4, worst
idx = lis.LateInteractionIndex(sum(p.shape[0] for p in pages), len(pages))
idx.add(pages, zero_pad_block=128)
q = [ru(16, 128), ru(100, 128)]
want = oracle.score_multi_vector_widened(q, pages)
for _ in range(3):                                                               # eager + capture, then graph replays
    v, i = idx.search(q, 7)
wv, wi = oracle.topk(want, 7)
assert torch.equal(i, wi) and (v - wv).abs().max().item() <= 1e-4
# K3: both rounding modes, dense and fused into the store
B, S, H = 5, 70, 256
hid = torch.randn(B, S, H, generator=g).to(torch.bfloat16).cuda()
w = (torch.randn(128, H, generator=g) / math.sqrt(H)).to(torch.bfloat16).cuda()
b = (0.1 * torch.randn(128, generator=g)).to(torch.bfloat16).cuda()
mask = (torch.rand(B, S, generator=g) > 0.3).long().cuda()
for mode in ("reference", "f32"):
    dense = lis.project_normalize(hid, w, b, mask, round_mode=mode)
    ix2 = lis.LateInteractionIndex(B * S, B)
    ix2.add_from_hidden(hid, w, b, mask, round_mode=mode)
    assert torch.equal(ix2.read_rows(0, ix2.num_rows), dense.cpu()[mask.cpu().bool()])
    ix2.close()
# fp32 planes
pf = [unit(torch.randn(n, 128, generator=g)) for n in (40, 300, 7)]
qf = [unit(torch.randn(16, 128, generator=g))]
assert (lis.score_multi_vector(qf, [p.cuda() for p in pf]) - oracle.score_multi_vector(qf, pf)).abs().max().item() <= 1e-4
idx.close()
torch.cuda.synchronize()
print(f"sanitize case ok (max abs err {worst:.2e})")
