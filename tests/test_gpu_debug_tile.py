"""First thing to run on a GPU: the raw TMA -> UMMA -> TMEM path against a plain matmul.
If this fails, everything downstream is noise; the printed pattern says which descriptor is wrong."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tile_n,a_in_tmem", [(128, 0), (256, 0), (128, 1), (192, 1)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_sim_tile_matches_matmul(lis, tile_n, a_in_tmem, dtype):
    from importlib import import_module

    N = import_module("multi-modal_colpali_b200._native")
    lib = N.load()
    g = torch.Generator().manual_seed(7)
    q = torch.randn(100, 128, generator=g).to(dtype).cuda()      # 28 rows short of a full M tile
    p = torch.randn(tile_n + 40, 128, generator=g).to(dtype).cuda()
    out = torch.full((128, tile_n), float("nan"), dtype=torch.float32, device="cuda")
    rc = lib.lis_debug_sim_tile(q.data_ptr(), q.shape[0], p.data_ptr(), p.shape[0], 0 if dtype == torch.bfloat16 else 1,
                                tile_n, a_in_tmem, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    N.check(rc)
    torch.cuda.synchronize()
    qf = torch.zeros(128, 128, device="cuda")
    qf[:100] = q.float()                                          # rows past the end must read as zero
    q = qf
    ref = q @ p[:tile_n].float().T
    err = (out - ref).abs()
    if not (err.max() < 1e-3):
        bad = (err > 1e-3)
        print("max err", err.max().item(), "bad fraction", bad.float().mean().item())
        print("bad rows", bad.any(1).nonzero().flatten()[:16].tolist(), "bad cols", bad.any(0).nonzero().flatten()[:16].tolist())
        print("out[0,:8]", out[0, :8].tolist(), "ref[0,:8]", ref[0, :8].tolist())
        # does the output match a permutation of K chunks / rows?  quick probes
        for name, alt in {
            "first 64 k only": q.float()[:, :64] @ p[:tile_n].float()[:, :64].T,
            "last 64 k only": q.float()[:, 64:] @ p[:tile_n].float()[:, 64:].T,
            "first 16 k only": q.float()[:, :16] @ p[:tile_n].float()[:, :16].T,
            "transposed": (q.float() @ p[:tile_n].float().T).T[:128, :tile_n] if tile_n == 128 else ref,
        }.items():
            print(name, (out - alt).abs().max().item())
    assert err.max() < 1e-3


@pytest.mark.parametrize("n_mt", [3, 4])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_sim_pair_matches_matmul(lis, n_mt, dtype):
    """CTA-pair form: M = 256 instructions (one query tile per CTA) and, for n_mt = 3, the final M = 128
    instruction whose accumulator is laid out 64 rows x two column halves per CTA."""
    from importlib import import_module

    N = import_module("multi-modal_colpali_b200._native")
    lib = N.load()
    g = torch.Generator().manual_seed(11)
    rows = n_mt * 128 - 28
    q = torch.randn(rows, 128, generator=g).to(dtype).cuda()
    p = torch.randn(256 + 40, 128, generator=g).to(dtype).cuda()
    out = torch.full((n_mt * 128, 256), float("nan"), dtype=torch.float32, device="cuda")
    rc = lib.lis_debug_sim_pair(q.data_ptr(), rows, p.data_ptr(), p.shape[0], 0 if dtype == torch.bfloat16 else 1,
                                n_mt, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    N.check(rc)
    torch.cuda.synchronize()
    qf = torch.zeros(n_mt * 128, 128, device="cuda")
    qf[:rows] = q.float()
    ref = qf @ p[:256].float().T
    err = (out - ref).abs()
    if not (err.max() < 1e-3):
        bad = (err > 1e-3) | err.isnan()
        print("max err", err.max().item(), "bad fraction", bad.float().mean().item())
        for t in range(n_mt):
            blk = bad[t * 128:(t + 1) * 128]
            print("tile", t, "bad rows", blk.any(1).nonzero().flatten()[:8].tolist(), "...", int(blk.any(1).sum()),
                  "bad cols", blk.any(0).nonzero().flatten()[:8].tolist(), "...", int(blk.any(0).sum()))
    assert err.max() < 1e-3
