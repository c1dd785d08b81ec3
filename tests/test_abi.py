"""The C-ABI library loads without a GPU and exports every symbol include/lis.h declares;
host-only entry points (query planning, argument validation) behave as documented."""
import ctypes as C
import importlib
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def native():
    return importlib.import_module("multi-modal_colpali_b200._native")


def declared_symbols():
    text = (ROOT / "include" / "lis.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lis_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(native):
    lib = native.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"liblis.so does not export {n}"
        assert n in native.SIGNATURES, f"_native.py does not bind {n}"
    assert set(native.SIGNATURES) == set(names)
    assert lib.lis_abi_version() == 2


def test_no_torch_types_in_abi():
    """Signatures are plain C: strip comments, then no C++/torch type may remain."""
    text = (ROOT / "include" / "lis.h").read_text()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for banned in ("torch", "at::", "std::", "Tensor", "template", "class "):
        assert banned not in code, banned


def ref_plan(lens):
    """Independent restatement of the segmentation rule."""
    segs, row = [], 0
    for q, n in enumerate(lens):
        while n > 0:
            take = min(n, 64 - row % 64)     # cut at multiples of 64 packed rows (half an M tile)
            segs.append((q, row, row + take))
            row += take
            n -= take
    return segs, (row + 127) // 128


@pytest.mark.parametrize("lens", [[16], [20] * 32, [32] * 1024, [100, 100, 300, 0, 5], [128, 128], [1] * 300, [0, 0, 7],
                                  [0, 100], [0, 30], [30, 0], [64, 0, 64]])
def test_plan_queries(lis, lens):
    plan = lis.plan_queries(lens)
    segs, tiles = ref_plan(lens)
    assert plan.n_seg == len(segs) and plan.n_mtiles == tiles and plan.nq == len(lens)
    assert list(zip(plan.seg_query.tolist(), plan.seg_lo.tolist(), plan.seg_hi.tolist())) == segs
    for t in range(tiles):
        inside = [s for s in range(plan.n_seg) if plan.seg_lo[s] // 128 == t]
        assert list(range(plan.mt_seg[t], plan.mt_seg[t + 1])) == inside
    for s in range(plan.n_seg):      # a segment never straddles half an M tile (hence never an M tile)
        assert plan.seg_lo[s] // 64 == (plan.seg_hi[s] - 1) // 64
    for q in range(len(lens)):
        mine = list(range(plan.seg_first[q], plan.seg_first[q + 1]))
        assert sum(plan.seg_hi[s] - plan.seg_lo[s] for s in mine) == lens[q]
    # direct == segment s IS query s: nothing cut and nothing empty (lens [0, 100] give 2 segments for 2 queries)
    assert plan.direct == (len(segs) == len(lens) and all(n > 0 for n in lens))


def test_plan_queries_errors(native):
    lib = native.load()
    lens = np.asarray([4, -1], np.int32)
    n_mt = C.c_int64()
    rc = lib.lis_plan_queries(lens.ctypes.data, 2, 0, None, None, None, 0, None, C.byref(n_mt))
    assert rc == native.LIS_E_INVALID
    assert "negative length" in native.last_error()
    with pytest.raises(ValueError):
        native.check(rc)
    # capacity too small
    lens = np.asarray([200], np.int32)
    buf = np.zeros(1, np.int32)
    rc = lib.lis_plan_queries(lens.ctypes.data, 1, 1, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, 8,
                              np.zeros(8, np.int32).ctypes.data, C.byref(n_mt))
    assert rc == native.LIS_E_INVALID and "capacity" in native.last_error()


def test_argument_validation_without_gpu(native):
    lib = native.load()
    assert lib.lis_set_tuning(64, 0, 0, 0, 0) == native.LIS_E_INVALID
    assert lib.lis_set_tuning(256, 4, 0, 0, 0) == native.LIS_E_INVALID
    assert lib.lis_set_tuning(0, 0, 0, 0, 0) == 0
    assert lib.lis_maxsim_scores(None, 0, None, None, None, 0, 0, None, 0, None, None, 0, 0, 0, None, 0, None) \
        == native.LIS_E_INVALID
    assert "null pointer" in native.last_error()
    assert lib.lis_topk(None, 0, 1, 1, None, 0, 5000, None, None, None, 0, None) == native.LIS_E_INVALID
    assert lib.lis_index_num_pages(None) == 0


def test_new_entry_points_validate_without_gpu(native):
    """ABI v2 additions: argument validation happens before any CUDA call."""
    lib = native.load()
    assert lib.lis_index_search_sharded(None, None, None, 0, None, None, None, 0, 0, None, 0, 0, 1, None, None, None) \
        == native.LIS_E_INVALID
    assert lib.lis_comm_unique_id(None, 0) == native.LIS_E_INVALID
    assert lib.lis_comm_rank(None) == 0 and lib.lis_comm_world(None) == 1
    assert lib.lis_stream_scores(None, 0, None, None, None, 0, 0, None, None, 0, None, None, 0, 0, 0, None, 0, 0, 0, None) \
        == native.LIS_E_INVALID
    assert lib.lis_index_add_projected(None, None, 0, 0, 0, None, None, None, 1, 0, None, None) == native.LIS_E_INVALID
    assert lib.lis_project_normalize(None, 1, 64, None, None, None, 0, 0, 7, None, None, None) == native.LIS_E_INVALID
    assert "round_mode" in native.last_error()
    assert lib.lis_nccl_version() >= 0
    lib.lis_stream_release()


def test_pass_costs_drive_the_planner(native):
    """lis_set_pass_costs (host-only): the plan follows the installed table; NULL restores the built-in one."""
    lib = native.load()
    buf = (C.c_int32 * 64)()
    plan = lambda n: [buf[i] for i in range(lib.lis_maxsim_pass_plan(n, buf, 64))]
    default7 = plan(7)
    assert default7 == [-7]
    single = np.asarray([0, 1.0, 2.0, 3.0], np.float32)
    pair = np.asarray([0, 0] + [100.0] * 9, np.float32)          # pairs made expensive: everything runs on single CTAs
    try:
        native.check(lib.lis_set_pass_costs(single.ctypes.data, pair.ctypes.data))
        assert all(x > 0 for x in plan(7)) and sum(plan(7)) == 7
        bad = np.asarray([0, -1.0, 2.0, 3.0], np.float32)
        assert lib.lis_set_pass_costs(bad.ctypes.data, None) == native.LIS_E_INVALID
    finally:
        native.check(lib.lis_set_pass_costs(None, None))
    assert plan(7) == default7


def test_topk_workspace_sizes(native):
    lib = native.load()
    assert lib.lis_topk_workspace_bytes(1, 100, 10) == 256           # one chunk: no scratch
    a = lib.lis_topk_workspace_bytes(1, 500_000, 10)
    b = lib.lis_topk_workspace_bytes(32, 500_000, 10)
    c = lib.lis_topk_workspace_bytes(1, 500_000, 100)
    assert 0 < a < b and a < c
    assert lib.lis_topk_workspace_bytes(1, 100, 5000) == 0           # k out of range


@pytest.mark.parametrize("n_tiles", list(range(1, 40)) + [64, 255, 256, 1000])
def test_pass_plan_covers_every_tile_once(native, n_tiles):
    """lis_maxsim_pass_plan (host-only): the passes add up to the tile count and only instantiated shapes appear."""
    import ctypes as C

    lib = native.load()
    buf = (C.c_int32 * 1024)()
    try:
        for tun, allowed_pair, allowed_single in (((0, 0, 0, 0, 0), {2, 3, 4, 5, 6, 7, 8, 9, 10}, {1, 2, 3}),
                                                  ((0, 0, 0, 0, 3), {2, 3, 4, 5, 6}, {1}),
                                                  ((0, 10, 0, 0, 3), {2, 3, 4, 5, 6, 7, 8, 9, 10}, {1}),
                                                  ((0, 4, 0, 0, 0), {2, 3, 4}, {1, 2, 3}),
                                                  ((0, 0, 0, 0, 1), set(), {1, 2, 3}),
                                                  ((128, 5, 0, 0, 1), set(), {1, 2, 3, 4, 5})):
            native.check(lib.lis_set_tuning(*tun))
            n = lib.lis_maxsim_pass_plan(n_tiles, buf, 1024)
            assert n > 0
            passes = [buf[i] for i in range(n)]
            assert sum(abs(x) for x in passes) == n_tiles, (tun, passes)
            for x in passes:
                assert (-x in allowed_pair) if x < 0 else (x in allowed_single), (tun, passes)
            assert lib.lis_maxsim_pass_plan(n_tiles, None, 0) == n
    finally:
        lib.lis_set_tuning(0, 0, 0, 0, 0)
    # auto: one CTA per SM up to 2 tiles, CTA pairs from 3 on
    n = lib.lis_maxsim_pass_plan(n_tiles, buf, 1024)
    if n_tiles <= 2:
        assert [buf[i] for i in range(n)] == [n_tiles]
    elif n_tiles in (3, 4, 5, 6, 7, 8, 9, 10):
        assert [buf[i] for i in range(n)] == [-n_tiles]
