"""The reference's own arithmetic run on the SAME B200 through torch (cuBLAS einsum + max + sum, 128 x 128 blocks --
the algorithm of colpali-engine's score_multi_vector, restated inline; SURVEY.md section 8d asks for this as the
"existing Blackwell kernel" baseline), next to the fused kernel, on a slice of BASELINE configs[1]:
32 queries x 20 tokens vs P pages x 1030 tokens, bf16.  (a) corpus resident in HBM, (b) corpus on the host like the
reference keeps it (05_experiment02.py:213-214: every 128-page block is copied to the device per call)."""
import importlib, json, sys, time
from pathlib import Path
import torch
from torch.nn.utils.rnn import pad_sequence

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
dev = torch.device("cuda", 0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def torch_route(qs, ps, device, batch_size=128):
    rows = []
    for i in range(0, len(qs), batch_size):
        qb = pad_sequence(list(qs[i:i + batch_size]), batch_first=True, padding_value=0).to(device)
        blocks = []
        for j in range(0, len(ps), batch_size):
            pb = pad_sequence(list(ps[j:j + batch_size]), batch_first=True, padding_value=0).to(device)
            blocks.append(torch.einsum("bnd,csd->bcns", qb, pb).max(dim=3)[0].sum(dim=2))
        rows.append(torch.cat(blocks, dim=1).cpu())
    return torch.cat(rows, dim=0).to(torch.float32)


g = torch.Generator().manual_seed(1002)
q = unit(torch.randn(32, 20, 128, generator=g)).to(torch.bfloat16)
p = unit(torch.randn(P, 1030, 128, generator=g)).to(torch.bfloat16)
q_dev, p_dev = q.to(dev), p.to(dev)
out = {"pages": P, "queries": "32 x 20 tokens", "dtype": "bf16"}
for name, qq, pp in (("torch_resident", q_dev, p_dev), ("torch_host_corpus", q_dev, p)):
    torch_route(qq, pp[:256], dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ref = torch_route(qq, pp, dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[name + "_ms"] = dt * 1e3
    out[name + "_pairs_per_s"] = 32 * P / dt
for _ in range(3):
    got = lis.score_multi_vector(q_dev, p_dev, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    got = lis.score_multi_vector(q_dev, p_dev, device=dev)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
out["fused_ms"] = dt * 1e3
out["fused_pairs_per_s"] = 32 * P / dt
out["speedup_vs_torch_resident"] = out["torch_resident_ms"] / out["fused_ms"]
out["speedup_vs_torch_host_corpus"] = out["torch_host_corpus_ms"] / out["fused_ms"]
diff = (got - ref).abs()
out["max_abs_diff_vs_torch_gpu_bf16"] = diff.max().item()
out["share_bit_identical"] = (diff == 0).float().mean().item()
print(json.dumps(out), flush=True)
(ROOT / "gpurun_out" / "torch_route.json").write_text(json.dumps(out) + "\n")
