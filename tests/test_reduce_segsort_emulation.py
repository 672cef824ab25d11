"""CPU suite: the index logic of k_reduce_segsort (in-row column sort inside the reduce pass,
spsparse_b200/csrc/reduce_segsort.cuh) emulated tile by tile at small tile sizes and checked against a plain stable
sort + fold: rows straddling tiles, rows reaching the window's edges, duplicate runs that cross a tile's end, the
array's ends, and the count of entries in rows longer than the limit (tools/emulate_reduce_segsort.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import emulate_reduce_segsort as emu  # noqa: E402


def test_tiles_agree_with_plain_sort_and_fold():
    rng = np.random.default_rng(2026)
    for t in range(240):
        tile, seg_max = [(8, 3), (16, 4), (16, 8), (32, 5), (64, 16)][t % 5]
        emu.trial(rng, tile, seg_max, ["add", "replace", "leave"][t % 3])


def test_long_rows_are_counted_entry_for_entry():
    rng = np.random.default_rng(2027)
    seen = sum(emu.long_row_trial(rng, *[(8, 3), (16, 4), (32, 5)][t % 3]) for t in range(120))
    assert seen > 20


def test_rows_exactly_at_the_limit_and_aligned_with_tiles():
    # every row has exactly seg_max entries and tiles start on row starts (tile = 4 rows), then shifted by one entry
    for shift in (0, 1, 3):
        seg_max, tile = 4, 16
        rows = np.concatenate([np.zeros(shift, dtype=np.int64), np.repeat(np.arange(1, 40), seg_max)])
        rng = np.random.default_rng(shift)
        cols = rng.integers(0, 3, len(rows))
        keys = (rows << 10) | cols
        vals = rng.standard_normal(len(rows))
        got = emu.reduce_segsort_tiles(keys, vals, 10, tile, seg_max)
        want = emu.plain(keys, vals, 10)
        assert got[4] == 0
        for g, w in zip(got[:4], want):
            assert np.array_equal(g, w)
