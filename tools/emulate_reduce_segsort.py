"""Tile-by-tile emulation (pure Python) of k_reduce_segsort's index logic (spsparse_b200/csrc/reduce_segsort.cuh):
window, predecessor-count walk with the "cut" rule, placement into the sorted slots, head detection, left-to-right
folds that may continue past the tile.  The tile size and the row-length limit are parameters, so the boundary cases
(rows straddling tiles, rows touching the window's edges, runs of duplicates crossing a tile's end, the array's
ends) can be hammered at sizes a CPU finishes in seconds.  Checked against a plain stable sort + fold.

    python tools/emulate_reduce_segsort.py            # random trials, prints a summary
Used by tests/test_reduce_segsort_emulation.py.
"""
from __future__ import annotations

import numpy as np

SENT = -1  # "belongs to no row" (the kernel's ~0)


def reduce_segsort_tiles(keys, vals, bits_lo, tile, seg_max, policy="add"):
    """keys/vals: grouped by row (key >> bits_lo), insertion order inside a row.
    Returns (out_keys, out_vals, row_start, row_id, n_long) the way the kernel's tiles produce them."""
    n = len(keys)
    H = seg_max
    W = tile + 2 * H
    SORTED = tile + H + 1
    lo_mask = (1 << bits_lo) - 1
    out_k, out_v, row_start, row_id = [], [], [], []
    n_long = 0
    for base in range(0, n, tile):
        tile_n = min(tile, n - base)
        raw = [int(keys[g]) if 0 <= g < n else SENT for g in range(base - H, base - H + W)]
        v = [float(vals[g]) if 0 <= g < n else 0.0 for g in range(base - H, base - H + W)]
        skeys = [None] * SORTED
        svals = [None] * SORTED
        for q in range(W):
            key = raw[q]
            if key == SENT:
                continue
            row, col = key >> bits_lo, key & lo_mask
            b = f = before = 0
            cut = False
            while b < seg_max:
                if q < 1 + b:
                    cut = True
                    break
                kk = raw[q - 1 - b]
                if kk == SENT or (kk >> bits_lo) != row:
                    break
                before += (kk & lo_mask) <= col
                b += 1
            while f < seg_max:
                if q + 1 + f >= W:
                    cut = True
                    break
                kk = raw[q + 1 + f]
                if kk == SENT or (kk >> bits_lo) != row:
                    break
                before += (kk & lo_mask) < col
                f += 1
            is_long = b == seg_max or f == seg_max or b + f + 1 > seg_max
            if H <= q < H + tile:
                n_long += is_long
            place = q if (cut or is_long) else q - b + before
            if H <= place + 1 < H + SORTED:
                L = place + 1 - H
                assert skeys[L] is None, "two entries placed in one slot"
                skeys[L], svals[L] = key, v[q]
        # every in-array position of the sorted range must have been filled (slot 0 is unused for the first tile)
        for L in range(SORTED):
            g = base - 1 + L
            if 0 <= g < n:
                assert skeys[L] is not None, f"slot {L} of tile at {base} left empty"
        # ---- the reduce part, thread by thread is not needed: a sequential scan over the tile is the same thing ----
        j = 0
        while j < tile_n:
            L = j + 1
            key = skeys[L]
            head = (base + j == 0) or key != skeys[L - 1]
            if not head:
                j += 1
                continue
            rhead = (base + j == 0) or (key >> bits_lo) != (skeys[L - 1] >> bits_lo)
            acc = svals[L]
            q = base + j + 1
            while q < n:
                Lq = q - base + 1
                if Lq >= SORTED:
                    break
                if skeys[Lq] != key:
                    break
                if policy == "add":
                    acc = acc + svals[Lq]
                elif policy == "replace":
                    acc = svals[Lq]
                q += 1
            if rhead:
                row_start.append(len(out_k))
                row_id.append(key >> bits_lo)
            out_k.append(key)
            out_v.append(acc)
            j += 1
    row_start.append(len(out_k))
    return np.array(out_k, dtype=np.int64), np.array(out_v), np.array(row_start), np.array(row_id), n_long


def plain(keys, vals, bits_lo, policy="add"):
    order = np.argsort(keys, kind="stable")
    k, v = keys[order], vals[order]
    out_k, out_v, row_start, row_id = [], [], [], []
    for i in range(len(k)):
        if i and k[i] == k[i - 1]:
            if policy == "add":
                out_v[-1] = out_v[-1] + v[i]
            elif policy == "replace":
                out_v[-1] = v[i]
            continue
        if not i or (k[i] >> bits_lo) != (k[i - 1] >> bits_lo):
            row_start.append(len(out_k))
            row_id.append(int(k[i]) >> bits_lo)
        out_k.append(int(k[i]))
        out_v.append(float(v[i]))
    row_start.append(len(out_k))
    return np.array(out_k, dtype=np.int64), np.array(out_v), np.array(row_start), np.array(row_id)


def random_case(rng, n, nrows, bits_lo, ncols, max_row=None):
    """Row-grouped input: rows ascending, columns in random (insertion) order with duplicates."""
    rows = np.sort(rng.integers(0, nrows, n))
    if max_row is not None:  # cap the row lengths
        keep = np.ones(n, dtype=bool)
        start = 0
        for i in range(1, n + 1):
            if i == n or rows[i] != rows[start]:
                if i - start > max_row:
                    keep[start + max_row:i] = False
                start = i
        rows = rows[keep]
        n = len(rows)
    cols = rng.integers(0, ncols, n)
    keys = (rows.astype(np.int64) << bits_lo) | cols
    vals = rng.standard_normal(n)
    return keys, vals


def trial(rng, tile, seg_max, policy):
    n = int(rng.integers(0, 12 * tile))
    nrows = max(1, int(n / rng.choice([1, 2, 3, seg_max - 1, seg_max])))
    bits_lo = 10
    keys, vals = random_case(rng, n, nrows, bits_lo, int(rng.choice([3, 8, 1000])), max_row=seg_max)
    got = reduce_segsort_tiles(keys, vals, bits_lo, tile, seg_max, policy)
    want = plain(keys, vals, bits_lo, policy)
    assert got[4] == 0, "rows were capped at seg_max, none may be reported long"
    for g, w, name in zip(got[:4], want, ("keys", "vals", "row_start", "row_id")):
        assert np.array_equal(g, w), f"{name} differ (n={len(keys)}, tile={tile}, seg_max={seg_max}, policy={policy})"
    return len(keys)


def long_row_trial(rng, tile, seg_max):
    """Uncapped rows: whenever a row is longer than seg_max the kernel must say so (the host then discards the output)."""
    n = int(rng.integers(1, 8 * tile))
    keys, vals = random_case(rng, n, max(1, n // (seg_max // 2 + 1)), 10, 50)
    rows = keys >> 10
    _, counts = np.unique(rows, return_counts=True)
    has_long = bool((counts > seg_max).any())
    got = reduce_segsort_tiles(keys, vals, 10, tile, seg_max)
    if has_long:
        assert got[4] > 0, "a long row went unnoticed"
        assert got[4] == int(counts[counts > seg_max].sum()), "every entry of a long row is counted exactly once"
    else:
        assert got[4] == 0
        want = plain(keys, vals, 10)
        for g, w in zip(got[:4], want):
            assert np.array_equal(g, w)
    return has_long


def main(trials=300, seed=7):
    rng = np.random.default_rng(seed)
    total = 0
    for t in range(trials):
        tile, seg_max = [(8, 3), (16, 4), (16, 8), (32, 5), (64, 16)][t % 5]
        total += trial(rng, tile, seg_max, ["add", "replace", "leave"][t % 3])
    longs = sum(long_row_trial(rng, *[(8, 3), (16, 4), (32, 5)][t % 3]) for t in range(trials // 2))
    print(f"{trials} capped trials ({total} entries) and {trials // 2} uncapped trials ({longs} with long rows): all agree")


if __name__ == "__main__":
    main()
