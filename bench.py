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
def host_threads() -> int:
    """All the host threads the reference arm may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which left
    the round-1 reference arm on ONE thread at N > 1: set the count explicitly instead of inheriting it."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, n)


def cpu_reference_run(steps: int, warmup: int, sample_pages, dtype=torch.bfloat16, budget_s: float = 20.0):
    """Time the restated score_multi_vector on the host cores.  ``sample_pages`` None -> sized from a probe call
    so that the whole run (warm-up + timed calls) is about ``budget_s`` seconds of CPU work on this box."""
    from oracle import maxsim_oracle as oracle  # the only place bench.py executes oracle/

    torch.set_num_threads(host_threads())
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
                  f"({torch.get_num_threads()} threads set explicitly, {os.cpu_count()} logical cpus)",
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
def percentile(sorted_vals, p):
    return sorted_vals[min(len(sorted_vals) - 1, int(len(sorted_vals) * p))]


def run_ours(args):
    import ctypes as C

    import numpy as np
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
    peaks = measured_peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    def timed(fn, steps, warmup):
        """W warm-up calls, then exactly `steps` calls between barrier + synchronize; CUDA events on the launch stream,
        one per step boundary, so the mean step and the whole region come from the SAME loop.  Max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        l0 = lib.lis_launch_count()
        ev[0].record()
        for i in range(steps):
            fn()
            ev[i + 1].record()
        barrier()
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
        return max_over_ranks(ev[0].elapsed_time(ev[steps])), lib.lis_launch_count() - l0, per

    # ---- corpus: ONE index per GPU.  The first `pages` pages are the configs[1] workload; the whole shard
    #      (`search_pages` pages, default 500 000 = 131.8 GB) is the configs[3] single-query search corpus. ----
    pages = args.pages
    search_pages = max(args.search_pages, pages)
    rows = pages * PAGE_TOK
    index = None
    while index is None:          # 500 000 pages are 131.8 GB: on a GPU with less free memory the search corpus shrinks (and says so)
        try:
            index = lis.LateInteractionIndex(search_pages * PAGE_TOK, search_pages, device=dev)
        except (MemoryError, RuntimeError):
            if search_pages <= pages:
                raise
            search_pages = max(pages, search_pages * 3 // 4)
    if world > 1:                 # every rank must hold the same number of pages (ids are rank * search_pages + local)
        t = torch.tensor([search_pages], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) != search_pages:
            search_pages = int(t.item())
            index.close()
            index = lis.LateInteractionIndex(search_pages * PAGE_TOK, search_pages, device=dev)
    index.fill_synthetic(search_pages, PAGE_TOK, seed=2002 + rank, id_base=rank * search_pages)
    whole = index._as_store()
    store = scoring.PageStore(whole.tokens[:rows], whole.offsets[:pages + 1], whole.clamp[:pages], pages)
    q_host = make_queries().pin_memory()
    q_dev = q_host.to(dev)
    pq = scoring.pack_queries(q_dev, dev)
    scores = torch.empty((NQ, pages), dtype=torch.float32, device=dev)

    # 1. device-resident: the hot path alone (K1 + the segment sums of the cut queries), inputs already in HBM
    def step_device():
        scoring.maxsim_scores_device(pq, store, "f32", out=scores)

    # 2. end to end through the public API: pinned host queries -> H2D -> kernels -> D2H of the [32, pages] result
    e2e_out = torch.empty((NQ, pages), dtype=torch.float32).pin_memory()
    corpus_view = store.tokens.view(pages, PAGE_TOK, DIM)

    def step_e2e():
        return lis.score_multi_vector(q_host, corpus_view, device=dev, round_mode="f32", out=e2e_out)

    with ClockSampler(local) as clk:
        ms_dev, launches, per_step = timed(step_device, args.steps, args.warmup)
        e2e_steps = max(3, args.steps // 2)
        ms_e2e, _, _ = timed(step_e2e, e2e_steps, 3)

    pairs_step = NQ * pages * world
    value = pairs_step * args.steps / (ms_dev * 1e-3)
    e2e_value = pairs_step * e2e_steps / (ms_e2e * 1e-3)

    # roofline of the dominant kernel (K1), from the per-step events of the timed region above
    plan_buf = (C.c_int32 * 16)()
    n_pass = lib.lis_maxsim_pass_plan(pq.plan.n_mtiles, plan_buf, 16)
    native.check(min(n_pass, 0))
    passes = [int(plan_buf[i]) for i in range(n_pass)]
    traffic, traffic_src = args.traffic, "--traffic"
    for name in ("k1_traffic_r2.json", "k1_traffic_r1.json"):
        tf = ROOT / "profiles" / name
        if traffic is None and tf.exists():
            traffic = json.loads(tf.read_text())["dram_bytes_per_page_token_row"] * rows * n_pass
            traffic_src = f"committed ncu capture profiles/{name} (dram bytes per page-token row) x rows x launches; not measured in this run"
    m_rows = NQ * QTOK
    flops = 2.0 * m_rows * DIM * rows                 # algorithmic: real query rows x real page rows
    bytes_alg = rows * DIM * 2.0 * n_pass             # page tokens, read once per launch (= per pass over the store)
    k1 = statistics.mean(per_step) * 1e-3
    ach_tf, ach_gbs = flops / k1 / 1e12, bytes_alg / k1 / 1e9
    t_mma, t_hbm = flops / (peaks["tf_burst"] * 1e12), bytes_alg / (peaks["hbm_gbs"] * 1e9)
    bound = "tensor" if t_mma >= t_hbm else "hbm"
    roofline = {
        "bound": bound, "achieved": ach_tf if bound == "tensor" else ach_gbs,
        "peak": peaks["tf_burst"] if bound == "tensor" else peaks["hbm_gbs"],
        "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
        "frac": (ach_tf / peaks["tf_burst"]) if bound == "tensor" else (ach_gbs / peaks["hbm_gbs"]),
        "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": f"{peaks['source']} (burst figure; sustained fraction alongside)",
        "kernel": " + ".join(f"lis::maxsim_pair_kernel[{-n} query tiles, CTA pairs]" if n < 0 else
                             f"lis::maxsim_kernel[{n} query tiles]" for n in passes),
        "kernel_ms": k1 * 1e3, "kernel_ms_source": "mean of the per-step CUDA events of the timed region (same loop as ms_per_step; "
                                                   "the step is K1 + a reduce_segments launch of < 0.5 %)",
        "launches_per_step": n_pass,
        "frac_of_sustained_tensor": ach_tf / peaks["tf_sustained"] if peaks["tf_sustained"] else None,
        "hbm_gbs": ach_gbs, "hbm_frac": ach_gbs / peaks["hbm_gbs"],
        "algorithmic": {"flops_per_step": flops, "bytes_per_step": bytes_alg,
                        "note": "2*query_rows*128 FLOP per page-token row (640 query rows); "
                                "256 B per page-token row per launch (the store is streamed once per launch)"},
    }

    # 3. tensor regime (BASELINE configs[4]: 1024 queries x 32 tokens), on a slice of the store sized to ~150 ms
    def tensor_regime():
        t_pages = min(pages, args.tensor_pages)
        t_rows = t_pages * PAGE_TOK
        g = torch.Generator().manual_seed(1005)
        qt = unit_rows(torch.randn(1024, 32, DIM, generator=g)).to(torch.bfloat16).to(dev)
        pqt = scoring.pack_queries(qt, dev)
        st = scoring.PageStore(whole.tokens[:t_rows], whole.offsets[:t_pages + 1], whole.clamp[:t_pages], t_pages)
        out = torch.empty((1024, t_pages), dtype=torch.float32, device=dev)
        ms, _, per = timed(lambda: scoring.maxsim_scores_device(pqt, st, "f32", out=out), 5, 3)
        fl = 2.0 * 1024 * 32 * DIM * t_rows
        tfs = fl / (statistics.mean(per) * 1e-3) / 1e12
        buf = (C.c_int32 * 64)()
        npass = lib.lis_maxsim_pass_plan(pqt.plan.n_mtiles, buf, 64)
        return {"what": f"BASELINE configs[4]: 1024 queries x 32 tokens vs {t_pages} pages x {PAGE_TOK} tokens per GPU "
                        f"(a slice of the 200 000-page corpus sized to ~150 ms), full score matrix",
                "ms": statistics.mean(per), "pairs_per_s": 1024 * t_pages * world / (statistics.mean(per) * 1e-3),
                "tflops": tfs, "frac_of_burst": tfs / peaks["tf_burst"],
                "frac_of_sustained": tfs / peaks["tf_sustained"] if peaks["tf_sustained"] else None,
                "passes": [int(buf[i]) for i in range(min(npass, 64))], "launches_per_step": int(npass)}

    tensor = tensor_regime() if not args.no_extra else None

    # 4. single-query top-10 search at BASELINE configs[3] scale: host in -> host out through ONE C call
    #    (K1 -> K2 -> [ncclAllGather -> merge] -> download, replayed as a CUDA graph)
    sharded = lis.ShardedIndex(index)
    q1_host = make_queries(1, 16, seed=1004)

    def latency(fn, iters, warm=5):
        lat = []
        for i in range(iters + warm):
            if world > 1 and i == warm:
                dist.barrier()
            t0 = time.perf_counter()
            fn()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = sorted(lat[warm:])
        latency.fastest_rank_p50 = min_over_ranks(percentile(lat, 0.5))
        return max_over_ranks(percentile(lat, 0.5)), max_over_ranks(percentile(lat, 0.95))

    p50, p95 = latency(lambda: sharded.search(q1_host, 10), args.search_iters)
    p50_fastest = latency.fastest_rank_p50
    floor_ms = search_pages * PAGE_TOK * DIM * 2.0 / (peaks["hbm_gbs"] * 1e9) * 1e3
    search = {"what": f"BASELINE configs[3] shape: 1 query x 16 tokens, top-10 over {search_pages * world} pages x {PAGE_TOK} tokens "
                      f"({search_pages} per GPU = {search_pages * PAGE_TOK * 256 / 1e9:.1f} GB), host in / host out, one call "
                      f"(lis_index_search_sharded: CUDA graph" + (", ncclAllGather + merge inside" if world > 1 else "") + ")",
              "p50_ms": p50, "p95_ms": p95, "iters": args.search_iters, "hbm_floor_ms": floor_ms,
              "roofline_frac": floor_ms / p50, "hbm_gbs": search_pages * PAGE_TOK * 256 / (p50 * 1e-3) / 1e9,
              "p50_ms_fastest_rank": p50_fastest,
              "timing": "host wall clock around the blocking call; p50 / p95 are the MAX over ranks (every rank waits for the slowest "
                        "GPU's shard inside the all-gather), p50_ms_fastest_rank the min"}
    # the exchange must not change the answer: merged sharded top-10 == host-side merge of every rank's own top-10
    sv, si = sharded.search(q1_host, 10)
    lv, li = index.search(q1_host, 10)
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (lv, li))
        allv = torch.cat([p[0] for p in parts], dim=1)[0]
        alli = torch.cat([p[1] for p in parts], dim=1)[0]
        order = sorted(range(allv.numel()), key=lambda j: (-float(allv[j]), int(alli[j])))[:10]
        want_i, want_v = alli[order], allv[order]
    else:
        want_i, want_v = li[0], lv[0]
    if not (torch.equal(si[0], want_i) and torch.equal(sv[0], want_v)):
        raise SystemExit(f"rank {rank}: sharded top-10 differs from the merge of the per-rank top-10")
    search["checked"] = "sharded top-10 == merge of every rank's local top-10 (ids and scores, bit-exact)"

    # 4b. the reference's own corpus size (05_experiment02: a few hundred pages): launch-bound regime
    small = lis.LateInteractionIndex(300 * PAGE_TOK, 300, device=dev)
    small.fill_synthetic(300, PAGE_TOK, seed=7, id_base=0)
    q10 = make_queries(10, 20, seed=1010)
    s50, s95 = latency(lambda: small.search(q10, 5), 200, warm=10)
    search_small = {"what": "reference-sized corpus: 10 queries x 20 tokens (one chunk of 05_experiment02.py:272), top-5 over 300 pages, "
                            "host in / host out, one call (CUDA graph)", "p50_ms": s50, "p95_ms": s95, "iters": 200}
    small.close()

    # 5. the reference's literal call: corpus on the HOST (05_experiment02.py:213-214), streamed in pinned chunks
    def host_corpus():
        hp = min(pages, args.host_pages)
        host = store.tokens[:hp * PAGE_TOK].view(hp, PAGE_TOK, DIM).cpu()           # pageable CPU tensor, like the reference's
        nbytes = host.numel() * 2
        pin = torch.empty(1 << 29, dtype=torch.uint8).pin_memory()                  # PCIe floor: pinned 512 MiB H2D copies
        pin.zero_()
        dst = torch.empty(1 << 29, dtype=torch.uint8, device=dev)
        best = 1e9
        for it in range(8):                                                         # 2 warm-ups, best of 6
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dst.copy_(pin, non_blocking=True); e1.record(); torch.cuda.synchronize()
            if it >= 2:
                best = min(best, e0.elapsed_time(e1))
        h2d_gbs = (1 << 29) / (best * 1e-3) / 1e9
        del pin, dst
        out = torch.empty((NQ, hp), dtype=torch.float32).pin_memory()
        fn = lambda: lis.score_multi_vector(q_host, host, device=dev, round_mode="f32", out=out)
        for _ in range(2):
            fn()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
        ms = statistics.median(ts)
        fn_pinned = None
        hpin = host.pin_memory()
        fnp = lambda: lis.score_multi_vector(q_host, hpin, device=dev, round_mode="f32", out=out)
        fnp(); tp = []
        for _ in range(5):
            t0 = time.perf_counter(); fnp(); tp.append((time.perf_counter() - t0) * 1e3)
        lib.lis_stream_release()
        floor = nbytes / (h2d_gbs * 1e9) * 1e3
        return {"what": f"score_multi_vector(pinned host queries, HOST-resident corpus of {hp} pages x {PAGE_TOK} tokens = "
                        f"{nbytes / 1e9:.2f} GB) -> CPU float32 [32, {hp}]: chunks of whole pages through a pinned double buffer, "
                        f"K1 overlapped with the next chunk's copy",
                "pageable_ms": ms, "pinned_ms": statistics.median(tp), "pcie_floor_ms": floor, "h2d_gbs_measured": h2d_gbs,
                "pageable_frac_of_floor": floor / ms, "pinned_frac_of_floor": floor / statistics.median(tp),
                "pairs_per_s": NQ * hp / (ms * 1e-3), "h2d_bytes_per_step": nbytes + NQ * QTOK * DIM * 2,
                "d2h_bytes_per_step": NQ * hp * 4,
                "note": "torch's GPU route of the reference loop with the same host corpus: 505 ms for 10 000 pages "
                        "(profiles/torch_gpu_route_r1.json)"}

    host = host_corpus() if (rank == 0 and not args.no_extra) else None
    barrier()

    # 6. BASELINE configs[2]: ColQwen2-like ragged pages (256..768 tokens), 1 M pages over the GPUs, top-100
    def ragged_leg():
        total_pages = args.ragged_pages
        a, b = lis.shard_range(total_pages, rank, world)
        lens = np.random.default_rng(3003).integers(256, 769, size=total_pages)[a:b].astype(np.int32)
        ridx = lis.LateInteractionIndex(int(lens.sum()), len(lens), device=dev)
        ridx.fill_synthetic(len(lens), lens, seed=2003 + rank, id_base=a)
        rs = lis.ShardedIndex(ridx)
        out = {"what": f"BASELINE configs[2]: {total_pages} ragged pages of 256..768 tokens ({total_pages // world} per GPU), "
                       f"top-100, host in / host out, one call"}
        gbytes = max_over_ranks(float(lens.sum()) * 256)
        for nq, name in ((1, "1_query_x_32_tokens"), (32, "32_queries_x_32_tokens")):
            qq = make_queries(nq, 32, seed=1003)
            p50r, p95r = latency(lambda: rs.search(qq, 100), 12, warm=3)
            rec = {"p50_ms": p50r, "p95_ms": p95r, "pairs_per_s": nq * total_pages / (p50r * 1e-3)}
            if nq == 1:
                rec["hbm_floor_ms"] = gbytes / (peaks["hbm_gbs"] * 1e9) * 1e3
                rec["roofline_frac"] = rec["hbm_floor_ms"] / p50r
            else:
                fl = 2.0 * nq * 32 * DIM * float(lens.sum())
                rec["tflops_per_gpu"] = fl / (p50r * 1e-3) / 1e12
                rec["frac_of_burst"] = rec["tflops_per_gpu"] / peaks["tf_burst"]
            out[name] = rec
        rs.close(); ridx.close()
        return out

    ragged = None
    if not args.no_extra:
        sharded.close(); index.close()
        del whole, store, corpus_view, scores
        torch.cuda.empty_cache()
        ragged = ragged_leg()
    barrier()

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
            "search": search,
            "search_reference_size": search_small,
        }
        if tensor is not None:
            line["tensor_regime"] = tensor
        if host is not None:
            line["e2e_host_corpus"] = host
        if ragged is not None:
            line["ragged_top100"] = ragged
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
    ap.add_argument("--search-pages", type=int, default=500_000,
                    help="pages per GPU of the single-query search leg (BASELINE configs[3]: 4 M pages over 8 GPUs)")
    ap.add_argument("--tensor-pages", type=int, default=24_000, help="pages of the configs[4] slice (tensor regime)")
    ap.add_argument("--host-pages", type=int, default=10_000, help="pages of the host-resident-corpus leg")
    ap.add_argument("--ragged-pages", type=int, default=1_000_000, help="total pages of the configs[2] leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the tensor-regime, host-corpus and ragged legs")
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
