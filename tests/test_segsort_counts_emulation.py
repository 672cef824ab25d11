"""CPU suite: the hand-over between the in-row sort and the reduce pass that needs no look-back (radix_sort.cuh:
k_segment_sort_walk / k_segment_sort count, per 256-entry tile of their OUTPUT, the run heads and row heads they placed there;
reduce_warp.cuh: k_reduce_warp<false> reads its place from the scanned counts), emulated at small tile sizes
(tools/emulate_segsort_counts.py) and checked against a plain stable sort + fold: every tile produces exactly what was counted
for it, with rows straddling tiles, duplicates across tile ends, every policy, KEEP_ALL."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import emulate_segsort_counts as emu  # noqa: E402


def test_counts_place_every_tile_without_lookback():
    rng = np.random.default_rng(4242)
    done = 0
    for t in range(360):
        wt, seg_max = [(8, 4), (16, 5), (16, 9), (32, 9), (64, 16)][t % 5]
        done += emu.trial(rng, wt, seg_max, ["add", "replace", "leave"][t % 3], keep_all=t % 7 == 0)
    assert done > 200      # (trials with a row over the limit are the library's fallback case, not this path)


def test_duplicates_across_tile_ends_and_rows_at_the_limit():
    # every row has exactly seg_max entries, all with the same few columns (long runs of duplicates), tiles cut rows anywhere
    for shift in (0, 1, 3, 5):
        wt, seg_max, bits_lo = 8, 6, 10
        rows = np.concatenate([np.zeros(shift, dtype=np.int64), np.repeat(np.arange(1, 30), seg_max)])
        rng = np.random.default_rng(shift)
        keys = (rows << bits_lo) | rng.integers(0, 2, len(rows))
        vals = rng.standard_normal(len(rows))
        sk, sv, heads, rheads, n_long = emu.in_row_sort_with_counts(keys, vals, bits_lo, wt, seg_max)
        assert n_long == 0 and heads.sum() <= 2 * 30 and rheads.sum() == len(np.unique(rows))
        for policy in ("add", "replace", "leave"):
            got = emu.reduce_tiles_without_lookback(sk, sv, bits_lo, wt, heads, rheads, policy)
            for g, w in zip(got, emu.plain(keys, vals, bits_lo, policy)):
                assert np.array_equal(g, w)


def test_long_rows_invalidate_the_counts():
    rng = np.random.default_rng(7)
    wt, seg_max, bits_lo = 16, 4, 10
    rows = np.sort(np.concatenate([np.full(9, 5), rng.integers(0, 40, 60)]))      # row 5 is over the limit
    keys = (rows.astype(np.int64) << bits_lo) | rng.integers(0, 30, len(rows))
    *_, n_long = emu.in_row_sort_with_counts(keys, rng.standard_normal(len(rows)), bits_lo, wt, seg_max)
    assert n_long >= 9
