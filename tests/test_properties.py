"""Property tests (hypothesis) of the host-side logic that decides what the kernels are asked to do: query segmentation,
zero-padding flags, shard ranges, batcher placement, pass plans.  CPU only."""
import ctypes as C
import importlib

import numpy as np
from hypothesis import given, settings, strategies as st

lis = importlib.import_module("multi-modal_colpali_b200")
native = importlib.import_module("multi-modal_colpali_b200._native")


@settings(max_examples=150, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=400), min_size=1, max_size=40))
def test_plan_queries_partitions_the_rows(lens):
    plan = lis.plan_queries(lens)
    total = sum(lens)
    assert plan.nq == len(lens) and plan.n_rows == total and plan.n_mtiles == (total + 127) // 128
    # segments tile the packed rows exactly, in order, never crossing a multiple of 64, each inside its query
    row = 0
    starts = np.concatenate([[0], np.cumsum(lens)])
    for s in range(plan.n_seg):
        lo, hi, q = int(plan.seg_lo[s]), int(plan.seg_hi[s]), int(plan.seg_query[s])
        assert lo == row and hi > lo and lo // 64 == (hi - 1) // 64
        assert starts[q] <= lo and hi <= starts[q + 1]
        row = hi
    assert row == total
    for q, n in enumerate(lens):
        a, b = int(plan.seg_first[q]), int(plan.seg_first[q + 1])
        assert sum(int(plan.seg_hi[s] - plan.seg_lo[s]) for s in range(a, b)) == n
        assert (a == b) == (n == 0)
    # direct <=> K1's output row s is query s: nothing cut, nothing empty
    assert plan.direct == (all(n > 0 for n in lens) and plan.n_seg == len(lens))
    for t in range(plan.n_mtiles):
        for s in range(int(plan.mt_seg[t]), int(plan.mt_seg[t + 1])):
            assert int(plan.seg_lo[s]) // 128 == t


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=50), min_size=1, max_size=300), st.integers(min_value=1, max_value=140))
def test_clamp_flags_mark_exactly_the_padded_pages(lens, bs):
    flags = lis.clamp_flags(lens, bs)
    for j in range(0, len(lens), bs):
        blk = lens[j:j + bs]
        assert flags[j:j + bs].tolist() == [int(n < max(blk)) for n in blk]


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=2000), min_size=0, max_size=200), st.integers(min_value=1, max_value=9))
def test_shard_ranges_are_contiguous_and_cover(lens, world):
    parts = lis.balanced_shard_ranges(lens, world)
    assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == len(lens)
    assert all(a <= b for a, b in parts) and all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    even = [lis.shard_range(len(lens), r, world) for r in range(world)]
    assert even[0][0] == 0 and even[-1][1] == len(lens)
    sizes = [b - a for a, b in even]
    assert max(sizes) - min(sizes) <= 1
    mine = lis.assign_shards([max(n, 1) for n in lens] or [1], world)
    assert mine[0][0] == 0 and mine[-1][1] == max(len(lens), 1)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.integers(min_value=1, max_value=200), min_size=1, max_size=30))
def test_batcher_placement_never_changes_a_querys_segmentation(lens):
    """Placed behind one another by QueryBatcher._placed_rows (padding included), every query is cut exactly where it is cut
    when it is searched alone -- the condition for bit-identical coalesced results."""
    rows, placed = 0, []
    for n in lens:
        end = lis.QueryBatcher._placed_rows(rows, n)
        placed.append((end - n, end))
        rows = end
    for (lo, hi), n in zip(placed, lens):
        alone = [c for c in range(64, n, 64)]                          # cuts of the query alone (offsets from its start)
        here = [c - lo for c in range((lo // 64 + 1) * 64, hi, 64)]     # cuts inside it where the batcher put it
        assert here == alone, (lens, lo, hi)


@settings(max_examples=60, deadline=None)
@given(st.integers(min_value=1, max_value=600))
def test_pass_plan_is_a_partition_into_available_shapes(n_tiles):
    lib = native.load()
    buf = (C.c_int32 * 1024)()
    n = lib.lis_maxsim_pass_plan(n_tiles, buf, 1024)
    passes = [buf[i] for i in range(n)]
    assert sum(abs(x) for x in passes) == n_tiles
    assert all((1 <= x <= 3) or (-10 <= x <= -2) for x in passes)
    assert n <= n_tiles // 10 + 2                                      # ten tiles per pass wherever possible
