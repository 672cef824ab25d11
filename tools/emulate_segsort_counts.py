"""Emulation (pure Python) of the head counts the in-row sort hands to the reduce pass, and of that reduce pass without a
look-back (spsparse_b200/csrc/radix_sort.cuh: k_segment_sort_walk / k_segment_sort; reduce_warp.cuh: k_reduce_warp<false>):

  * the in-row sort decides for every entry of the row-grouped array where it lands in the sorted array (row start + number of
    row mates that must precede it), whether it repeats an EARLIER entry of its row (then the reduce pass folds it into that
    one) and whether it is the first of its row in column order; a "warp" takes WT consecutive entries and adds the heads /
    row heads it placed to three counters -- the tile of the reduce pass its entries start in, the one before, the one after;
  * a scan of the tile counts gives every reduce tile its exclusive prefix (entries, rows);
  * every reduce tile then works alone: heads from key != predecessor, followers folded left to right, outputs written at
    prefix + local rank, row starts at row prefix + local row rank.

The tile size and the row-length limit are parameters, so rows straddling tiles, duplicates across tile ends, rows at the limit
and over it (which invalidate the counts: the library then runs the look-back kernel) are hammered at sizes a CPU finishes in
seconds.  Checked against a plain stable sort + fold.

    python tools/emulate_segsort_counts.py            # random trials, prints a summary
Used by tests/test_segsort_counts_emulation.py.
"""
from __future__ import annotations

import numpy as np


def in_row_sort_with_counts(keys, vals, bits_lo, wt, seg_max, keep_all=False):
    """keys/vals grouped by row (key >> bits_lo), insertion order inside a row.
    -> (sorted keys, sorted vals, head count per tile, row-head count per tile, entries in rows longer than seg_max)"""
    n = len(keys)
    lo_mask = (1 << bits_lo) - 1
    out_k, out_v = np.zeros(n, dtype=np.int64), np.zeros(n)
    tiles = (n + wt - 1) // wt
    heads, rheads = np.zeros(tiles + 2, dtype=np.int64), np.zeros(tiles + 2, dtype=np.int64)
    n_long = 0
    for wbase in range(0, n, wt):                      # one "warp": wt consecutive entries = one tile of the reduce pass
        cnt = {-1: [0, 0], 0: [0, 0], 1: [0, 0]}      # heads, row heads by destination tile: before / mine / after
        for g in range(wbase, min(wbase + wt, n)):
            row, col = int(keys[g]) >> bits_lo, int(keys[g]) & lo_mask
            b = f = before = same = 0
            while b < seg_max and g - 1 - b >= 0 and (int(keys[g - 1 - b]) >> bits_lo) == row:
                c = int(keys[g - 1 - b]) & lo_mask
                before += c <= col
                same += c == col
                b += 1
            while f < seg_max and g + 1 + f < n and (int(keys[g + 1 + f]) >> bits_lo) == row:
                before += (int(keys[g + 1 + f]) & lo_mask) < col
                f += 1
            is_long = b == seg_max or f == seg_max or b + f + 1 > seg_max
            dst, head, rhead = g, True, False
            if not is_long:
                dst = g - b + before
                head = keep_all or same == 0
                rhead = before == 0
            n_long += is_long
            out_k[dst], out_v[dst] = keys[g], vals[g]
            where = -1 if dst < wbase else (1 if dst >= wbase + wt else 0)
            assert abs(dst - g) < seg_max and (where == 0 or wbase // wt + where in range(tiles))
            cnt[where][0] += head
            cnt[where][1] += rhead
        t = wbase // wt
        for d in (-1, 0, 1):
            heads[t + d + 1] += cnt[d][0]             # (+1: slot 0 stands for "tile -1", never touched)
            rheads[t + d + 1] += cnt[d][1]
    assert heads[0] == 0 and heads[tiles + 1] == 0
    return out_k, out_v, heads[1:tiles + 1], rheads[1:tiles + 1], n_long


def reduce_tiles_without_lookback(keys, vals, bits_lo, wt, heads, rheads, policy="add", keep_all=False):
    """Every tile alone: its exclusive prefixes come from the scanned counts, nothing from its predecessors' work."""
    n = len(keys)
    excl_e = np.concatenate([[0], np.cumsum(heads)])
    excl_r = np.concatenate([[0], np.cumsum(rheads)])
    n_out, n_rows = int(excl_e[-1]), int(excl_r[-1])
    out_k, out_v = np.full(n_out, -1, dtype=np.int64), np.zeros(n_out)
    row_start, row_id = np.full(n_rows + 1, -1, dtype=np.int64), np.full(n_rows, -1, dtype=np.int64)
    for t, base in enumerate(range(0, n, wt)):
        slot, rslot = int(excl_e[t]), int(excl_r[t])
        for g in range(base, min(base + wt, n)):
            prev = int(keys[g - 1]) if g else None
            k = int(keys[g])
            head = keep_all or g == 0 or k != prev
            if not head:
                continue                                # folded by its run's head (possibly in an earlier tile)
            acc = float(vals[g])
            q = g + 1
            while not keep_all and q < n and int(keys[q]) == k:   # left to right, past the tile's end if need be
                acc = acc + float(vals[q]) if policy == "add" else (float(vals[q]) if policy == "replace" else acc)
                q += 1
            out_k[slot], out_v[slot] = k, acc
            if g == 0 or (k >> bits_lo) != (prev >> bits_lo):
                row_start[rslot], row_id[rslot] = slot, k >> bits_lo
                rslot += 1
            slot += 1
        assert slot == excl_e[t + 1] and rslot == excl_r[t + 1], "the tile produced what the in-row sort counted for it"
    row_start[n_rows] = n_out
    return out_k, out_v, row_start, row_id


def plain(keys, vals, bits_lo, policy="add", keep_all=False):
    o = np.argsort(keys, kind="stable")
    k, v = keys[o], vals[o]
    ok, ov = [], []
    for i in range(len(k)):
        if keep_all or i == 0 or k[i] != k[i - 1]:
            ok.append(int(k[i])); ov.append(float(v[i]))
        elif policy == "add":
            ov[-1] = ov[-1] + float(v[i])
        elif policy == "replace":
            ov[-1] = float(v[i])
    ok = np.array(ok, dtype=np.int64)
    rows = ok >> bits_lo
    starts = [i for i in range(len(ok)) if i == 0 or rows[i] != rows[i - 1]]
    return ok, np.array(ov), np.array(starts + [len(ok)], dtype=np.int64), rows[starts] if len(ok) else np.zeros(0, dtype=np.int64)


def grouped_input(rng, n, n_rows, n_cols, bits_lo):
    """Entries grouped by row, a random (insertion) order inside every row."""
    rows = np.sort(rng.integers(0, n_rows, n))
    cols = rng.integers(0, n_cols, n)
    return (rows.astype(np.int64) << bits_lo) | cols, rng.standard_normal(n)


def trial(rng, wt, seg_max, policy, keep_all=False):
    n = int(rng.integers(1, 6 * wt))
    bits_lo = 10
    max_rows = max(1, n // max(1, int(rng.integers(1, seg_max))))
    keys, vals = grouped_input(rng, n, max_rows, int(rng.integers(1, 40)), bits_lo)
    sk, sv, heads, rheads, n_long = in_row_sort_with_counts(keys, vals, bits_lo, wt, seg_max, keep_all)
    if n_long:
        return False                                    # counts are not valid: the library takes the look-back kernel
    got = reduce_tiles_without_lookback(sk, sv, bits_lo, wt, heads, rheads, policy, keep_all)
    want = plain(keys, vals, bits_lo, policy, keep_all)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    return True


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    done = sum(trial(rng, *[(8, 4), (16, 5), (32, 9)][t % 3], ["add", "replace", "leave"][t % 3], t % 7 == 0) for t in range(300))
    print(f"{done} of 300 trials had no long row and agree with a plain stable sort + fold")
