#!/usr/bin/env python
"""Headline benchmark: MaxSim query x page pairs/s (BASELINE.json metric) on BASELINE configs[1]
("ColPali bf16: 32 queries x 20 tokens vs 100k pages x 1030 tokens on 1 B200"), one corpus shard of
that size per GPU (weak scaling: pages shard naturally, no data-path collective for the score
matrix), plus the single-query top-10 search latency (local top-k -> one all-gather -> merge).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pages P]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

NQ, QTOK, PAGE_TOK, DIM = 32, 20, 1030, 128
DEFAULT_PAGES = 100_000
METRIC = "maxsim_query_page_pairs_per_s"
UNIT = "pairs/s"


def unit_rows(x):
    return x / x.norm(dim=-1, keepdim=True)


def make_queries(nq=NQ, qtok=QTOK, seed=1002):
    g = torch.Generator().manual_seed(seed)
    return unit_rows(torch.randn(nq, qtok, DIM, generator=g)).to(torch.bfloat16)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML, 100 ms period)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (a port of colpali-engine's score_multi_vector; the
# package itself is not installable here -- DESIGN.md) on the host cores, bounded sample.
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, sample_pages, dtype=torch.bfloat16, budget_s: float = 20.0):
    """Time the restated score_multi_vector on the host cores.  ``sample_pages`` None -> sized from a probe call
    so that the whole run (warm-up + timed calls) is about ``budget_s`` seconds of CPU work on this box."""
    from oracle import maxsim_oracle as oracle  # the only place bench.py executes oracle/

    q = make_queries().to(dtype)
    g = torch.Generator().manual_seed(2002)
    if sample_pages is None:
        probe = unit_rows(torch.randn(64, PAGE_TOK, DIM, generator=g)).to(dtype)
        oracle.score_multi_vector(q, probe, device="cpu")
        t0 = time.perf_counter()
        oracle.score_multi_vector(q, probe, device="cpu")
        per_page = (time.perf_counter() - t0) / 64
        sample_pages = int(budget_s / max(steps + warmup, 1) / max(per_page, 1e-9))
        sample_pages = max(128, min(sample_pages, 8192)) // 128 * 128     # whole 128-page blocks, <= 2.2 GB of bf16
    p = unit_rows(torch.randn(sample_pages, PAGE_TOK, DIM, generator=g)).to(dtype)
    for _ in range(warmup):
        oracle.score_multi_vector(q, p, device="cpu")
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        oracle.score_multi_vector(q, p, device="cpu")
        times.append(time.perf_counter() - t0)
    pairs = NQ * sample_pages
    total = sum(times)
    return {
        "value": pairs * steps / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
        "sample": f"{NQ} queries x {QTOK} tokens vs {sample_pages} pages x {PAGE_TOK} tokens, {str(dtype).split('.')[-1]}, "
                  f"{steps} timed calls ({total:.1f} s) of the restated score_multi_vector on CPU torch "
                  f"({torch.get_num_threads()} threads, {os.cpu_count()} logical cpus)",
        "ms_per_step": 1e3 * total / steps, "sample_pages": sample_pages,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference_run(args.steps, args.warmup, args.ref_pages, budget_s=120.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, res["sample_pages"], note="bounded sample of the same workload on host cores"),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, pages, note=None):
    cfg = {
        "workload": f"BASELINE configs[1]: ColPali bf16, {NQ} queries x {QTOK} tokens vs {pages} pages x {PAGE_TOK} tokens "
                    f"per GPU (128-d), full score matrix",
        "queries": NQ, "query_tokens": QTOK, "pages_per_gpu": pages, "page_tokens": PAGE_TOK, "dim": DIM,
        "sharding": f"pages x{args.gpus} (one shard of {pages} pages per GPU, no data-path collective)",
        "l2": f"inputs ({pages * PAGE_TOK * DIM * 2 / 1e9:.1f} GB/GPU) are larger than L2; no flush needed",
    }
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lis = importlib.import_module("multi-modal_colpali_b200")
    native = importlib.import_module("multi-modal_colpali_b200._native")
    lib = native.load()
    scoring = importlib.import_module("multi-modal_colpali_b200.scoring")

    pages = args.pages
    rows = pages * PAGE_TOK
    index = lis.LateInteractionIndex(rows, pages, device=dev)
    index.fill_synthetic(pages, PAGE_TOK, seed=2002 + rank, id_base=rank * pages)
    store = index._as_store()
    q_host = make_queries().pin_memory()
    q_dev = q_host.to(dev)
    pq = scoring.pack_queries(q_dev, dev)
    scores = torch.empty((NQ, pages), dtype=torch.float32, device=dev)
    sharded = lis.ShardedIndex(index)
    q1_host = make_queries(1, 16, seed=1004).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.lis_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), lib.lis_launch_count() - l0

    # 1. device-resident: the hot path alone (K1 + segment reduction), inputs already in HBM
    def step_device():
        scoring.maxsim_scores_device(pq, store, "f32", out=scores)

    # 2. end to end through the public API: host queries -> H2D -> kernels -> D2H of the result
    e2e_out = torch.empty((NQ, pages), dtype=torch.float32).pin_memory()

    def step_e2e():
        return lis.score_multi_vector(q_host, store.tokens.view(pages, PAGE_TOK, DIM), device=dev, round_mode="f32",
                                      out=e2e_out)

    # 3. single-query top-10 search (the latency half of the metric), incl. all-gather + merge
    def step_search():
        return sharded.search(q1_host, 10)

    with ClockSampler(local) as clk:
        ms_dev, launches = timed(step_device, args.steps, args.warmup)
        ms_e2e, _ = timed(step_e2e, max(3, args.steps // 2), 3)
    # kernel-only duration of K1 (dominant kernel) on its launch stream, for the roofline
    k1_ms = []
    for _ in range(max(5, args.steps)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_device()
        e1.record()
        torch.cuda.synchronize()
        k1_ms.append(e0.elapsed_time(e1))
    lat = []
    for i in range(args.search_iters + 5):
        t0 = time.perf_counter()
        step_search()
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = sorted(lat[5:])
    barrier()

    pairs_step = NQ * pages * world
    value = pairs_step * args.steps / (ms_dev * 1e-3)
    e2e_steps = max(3, args.steps // 2)
    e2e_value = pairs_step * e2e_steps / (ms_e2e * 1e-3)

    peaks = measured_peaks()
    # how K1 covers the 5 query M tiles: one entry per pass over the store (+n: n tiles on one CTA per SM, -n: CTA pairs)
    import ctypes as C
    plan_buf = (C.c_int32 * 16)()
    n_pass = lib.lis_maxsim_pass_plan(pq.plan.n_mtiles, plan_buf, 16)
    native.check(min(n_pass, 0))
    passes = [int(plan_buf[i]) for i in range(n_pass)]
    traffic = args.traffic
    tf = ROOT / "profiles" / "k1_traffic_r1.json"
    if traffic is None and tf.exists():
        # dram__bytes_read+write per page-token row from the committed ncu capture, x rows x launches
        traffic = json.loads(tf.read_text())["dram_bytes_per_page_token_row"] * rows * n_pass
    m_rows = NQ * QTOK
    flops = 2.0 * m_rows * DIM * rows                 # algorithmic: real query rows x real page rows
    bytes_alg = rows * DIM * 2.0 * n_pass             # page tokens, read once per launch (= per pass over the store)
    k1 = statistics.mean(k1_ms) * 1e-3
    ach_tf = flops / k1 / 1e12
    ach_gbs = bytes_alg / k1 / 1e9
    t_mma, t_hbm = flops / (peaks["tf_burst"] * 1e12), bytes_alg / (peaks["hbm_gbs"] * 1e9)
    bound = "tensor" if t_mma >= t_hbm else "hbm"
    roofline = {
        "bound": bound, "achieved": ach_tf if bound == "tensor" else ach_gbs,
        "peak": peaks["tf_burst"] if bound == "tensor" else peaks["hbm_gbs"],
        "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
        "frac": (ach_tf / peaks["tf_burst"]) if bound == "tensor" else (ach_gbs / peaks["hbm_gbs"]),
        "traffic": traffic, "peak_source": f"{peaks['source']} (burst; kernel timed alone)",
        "kernel": " + ".join(f"lis::maxsim_pair_kernel[{-n} query tiles, CTA pairs]" if n < 0 else
                             f"lis::maxsim_kernel[{n} query tiles]" for n in passes) + " (totals of all launches of a step)",
        "kernel_ms": k1 * 1e3, "launches_per_step": n_pass,
        "frac_of_sustained_tensor": ach_tf / peaks["tf_sustained"] if peaks["tf_sustained"] else None,
        "hbm_gbs": ach_gbs, "hbm_frac": ach_gbs / peaks["hbm_gbs"],
        "algorithmic": {"flops_per_step": flops, "bytes_per_step": bytes_alg,
                        "note": "2*query_rows*128 FLOP per page-token row (640 query rows); "
                                "256 B per page-token row per launch (the store is streamed once per launch)"},
    }

    if rank == 0:
        cpu = cpu_reference_run(3, 1, args.ref_pages, budget_s=20.0) if world == 1 and not args.no_cpu else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(args, pages),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": NQ * QTOK * DIM * 2 * world,
                    "d2h_bytes_per_step": NQ * pages * 4 * world, "ms_per_step": ms_e2e / e2e_steps,
                    "api": "score_multi_vector(pinned host queries, HBM-resident corpus, out=pinned host buffer) -> CPU float32 [32, pages]"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": roofline,
            "search": {"what": f"1 query x 16 tokens, top-10 over {pages * world} pages, host in / host out"
                               + (", all-gather + merge" if world > 1 else ""),
                       "p50_ms": lat[len(lat) // 2], "p95_ms": lat[int(len(lat) * 0.95)], "iters": len(lat),
                       "hbm_floor_ms": rows * DIM * 2.0 / (peaks["hbm_gbs"] * 1e9) * 1e3},
        }
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--pages", type=int, default=DEFAULT_PAGES, help="pages per GPU")
    ap.add_argument("--ref-pages", type=int, default=None,
                    help="pages in the bounded CPU sample (default: sized from a probe call, ~20 s / ~2 min of CPU work)")
    ap.add_argument("--search-iters", type=int, default=50)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--traffic", type=float, default=None,
                    help="dram bytes per K1 launch from the committed ncu capture (profiles/), if known")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
