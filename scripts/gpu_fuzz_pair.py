"""Randomised stress of the CTA-pair K1 against the single-CTA K1 (bit-identical) and the CPU oracle (1e-4, fp32
rounding mode): random query / page lengths (incl. empty and 1-token pages, >16 segments per tile, pages longer
than a tile), tile counts 2..23, CTA caps (long per-pair page ranges, page-table window refills), fp16 and bf16,
zero-padding block sizes.  Usage: gpu_fuzz_pair.py [cases] [seed]"""
import importlib, sys
from pathlib import Path
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
N = importlib.import_module("multi-modal_colpali_b200._native")
from oracle import maxsim_oracle as oracle   # checker (this script is a test driver, not product code)
lib = N.load()
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1234
g = torch.Generator().manual_seed(seed)


def unit(x):
    return x / x.norm(dim=-1, keepdim=True).clamp_min(1e-20)


def rint(lo, hi):
    return int(torch.randint(lo, hi + 1, (1,), generator=g))


bad = 0
for c in range(cases):
    dtype = torch.bfloat16 if rint(0, 3) else torch.float16
    n_tiles = rint(2, 23)
    rows = n_tiles * 128 - rint(0, 127)
    style = rint(0, 3)
    q_lens, left = [], rows
    while left > 0:
        n = {0: rint(1, 40), 1: rint(1, 3), 2: rint(60, 300), 3: 20}[style]
        n = min(n, left); q_lens.append(n); left -= n
    n_pages = rint(1, 900)
    pstyle = rint(0, 5)          # 4, 5: tile-aligned lengths with many empty pages (the page-close order cases)
    aligned = [0, 0, 0, 64, 128, 192, 256, 256, 320, 512, 1024]
    p_lens = [{0: rint(0, 60), 1: rint(200, 1100), 2: 1030, 3: rint(0, 3) * rint(0, 700), 4: aligned[rint(0, 10)],
               5: aligned[rint(0, 10)] + (rint(0, 9) == 0) * rint(1, 300)}[pstyle] for _ in range(n_pages)]
    bs = [128, 16, 7][rint(0, 2)]
    for j in range(0, n_pages, bs):          # the reference itself fails on a block of only empty pages (max over an empty dim)
        if max(p_lens[j:j + bs]) == 0:
            p_lens[j] = rint(1, 9)
    qs = [unit(torch.randn(n, 128, generator=g)).to(dtype) for n in q_lens]
    ps = [unit(torch.randn(n, 128, generator=g)).to(dtype) if n else torch.zeros(0, 128, dtype=dtype) for n in p_lens]
    ctas = [0, 0, 2, 6, 20][rint(0, 4)]
    grp = [0, 0, 4, 6, 10][rint(0, 4)]
    try:
        N.check(lib.lis_set_tuning(0, 0, ctas, 0, 1))
        single = lis.score_multi_vector(qs, ps, batch_size=bs, round_mode="f32")
        single16 = lis.score_multi_vector(qs, ps, batch_size=bs)
        N.check(lib.lis_set_tuning(0, grp, ctas, 0, 3))
        pair = lis.score_multi_vector(qs, ps, batch_size=bs, round_mode="f32")
        pair16 = lis.score_multi_vector(qs, ps, batch_size=bs)
        N.check(lib.lis_set_tuning(0, 0, ctas, 0, 0))
        auto = lis.score_multi_vector(qs, ps, batch_size=bs, round_mode="f32")
    finally:
        lib.lis_set_tuning(0, 0, 0, 0, 0)
    want = oracle.score_multi_vector_widened(qs, ps, batch_size=bs)
    err = (pair - want).abs().max().item()
    e_single = (single - want).abs().max().item()
    d_ps, d_ps16, d_as = (pair - single).abs().max().item(), (pair16 - single16).abs().max().item(), (auto - single).abs().max().item()
    ok = d_ps == 0 and d_ps16 == 0 and d_as == 0 and err <= 1e-4 and e_single <= 1e-4
    bad += 0 if ok else 1
    if not ok and e_single > 1e-4:
        wrong = (single - want).abs() > 1e-4
        pg = wrong.any(0).nonzero().flatten().tolist()
        qq = wrong.any(1).nonzero().flatten().tolist()
        import numpy as np
        off = np.concatenate([[0], np.cumsum(p_lens)])
        grid = min(148, ctas) if ctas else 148
        grid = min(grid, max(n_pages, 1))
        per = -(-int(off[-1]) // grid)
        print(f"   wrong pages {pg[:24]} (n={len(pg)}) wrong queries {qq[:12]} (n={len(qq)}, of {len(q_lens)})")
        for pgi in pg[:6]:
            lo = max(0, pgi - 3)
            print(f"   page {pgi}: len {p_lens[pgi]} off {int(off[pgi])} cta {int(off[pgi]) // per} neighbours lens {p_lens[lo:pgi + 4]} "
                  f"got {single[qq[0], pgi].item():.4f} want {want[qq[0], pgi].item():.4f}")
        print(f"   per-CTA rows {per}, boundaries (first pages) {[int(np.searchsorted(off, b * per, side='left')) for b in range(min(grid, 8))]}")
    print(f"case {c}: tiles={n_tiles} q={len(q_lens)}(style {style}) pages={n_pages}(style {pstyle}, {sum(p_lens)} rows) {str(dtype)[6:]} bs={bs} "
          f"ctas={ctas} grp={grp} err={err:.2e} {'ok' if ok else f'MISMATCH single-vs-oracle {e_single:.2e} pair-single {d_ps:.2e} (16: {d_ps16:.2e}) auto-single {d_as:.2e} n_bad {int((pair != single).sum())}'}", flush=True)
print("FUZZ", "PASS" if bad == 0 else f"FAIL ({bad})")
sys.exit(0 if bad == 0 else 1)
