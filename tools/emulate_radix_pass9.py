"""Emulation (pure Python) of the per-tile arithmetic of k_radix_pass9 (spsparse_b200/csrc/radix_sort9.cuh): warp-striped
items, per-warp 16-bit digit counters packed two to a 32-bit word, the digit-pair owner threads (totals, start inside
the tile, running offsets of each warp written back into the 16-bit halves), staging positions, per-digit global bases
from the pass histogram plus the preceding tiles' counts, final destinations.  The result must be the stable sort of the
keys by their 9-bit digit.  The look-back itself (a chain of relaxed loads) is replaced by the sum it computes.

    python tools/emulate_radix_pass9.py
Used by tests/test_radix_pass9_emulation.py.
"""
from __future__ import annotations

import numpy as np

THREADS, IPT, WARPS, RADIX = 256, 16, 8, 512
TILE = THREADS * IPT


def pass9(keys, shift):
    n = len(keys)
    digit = lambda k: (int(k) >> shift) & (RADIX - 1)  # noqa: E731
    hist = np.zeros(RADIX, dtype=np.int64)
    for k in keys:
        hist[digit(k)] += 1
    # k_bucket_starts9: thread t owns digits 2t, 2t+1
    bucket_start = np.zeros(RADIX, dtype=np.int64)
    run = 0
    for t in range(RADIX // 2):
        c0, c1 = hist[2 * t], hist[2 * t + 1]
        incl = run + c0 + c1
        bucket_start[2 * t] = incl - c0 - c1
        bucket_start[2 * t + 1] = incl - c1
        run = incl
    out = [None] * n
    prev_tiles = np.zeros(RADIX, dtype=np.int64)  # what the look-back adds up: counts of the preceding tiles, per digit
    for tile_base in range(0, n, TILE):
        cnt2 = np.zeros((WARPS, RADIX // 2), dtype=np.uint32)   # two 16-bit counters per word
        item = {}   # (warp, k, lane) -> (key, pos)
        for warp in range(WARPS):
            for k in range(IPT):
                lanes = []
                for lane in range(32):
                    i = tile_base + warp * 32 * IPT + k * 32 + lane
                    if i < n:
                        lanes.append((lane, keys[i]))
                # ballot multi-split: lanes with the same digit, in lane order; the lowest lane bumps the counter
                seen = {}
                for lane, key in lanes:
                    seen.setdefault(digit(key), []).append(lane)
                for d, ls in seen.items():
                    word, half = d >> 1, d & 1
                    before = (int(cnt2[warp, word]) >> (16 * half)) & 0xFFFF
                    new = before + len(ls)
                    assert new <= 0xFFFF
                    cnt2[warp, word] = (int(cnt2[warp, word]) & ~(0xFFFF << (16 * half)) & 0xFFFFFFFF) | (new << (16 * half))
                    for r, lane in enumerate(ls):
                        item[(warp, k, lane)] = before + r
        # ---- digit-pair owners
        tot = np.zeros(RADIX, dtype=np.int64)
        lstart = np.zeros(RADIX, dtype=np.int64)
        run = 0
        for t in range(THREADS):
            tot0 = sum(int(cnt2[w, t]) & 0xFFFF for w in range(WARPS))
            tot1 = sum(int(cnt2[w, t]) >> 16 for w in range(WARPS))
            tot[2 * t], tot[2 * t + 1] = tot0, tot1
            lstart[2 * t] = run
            lstart[2 * t + 1] = run + tot0
            run0, run1 = run, run + tot0
            for w in range(WARPS):
                c = int(cnt2[w, t])
                assert run0 <= 0xFFFF and run1 <= 0xFFFF
                cnt2[w, t] = (run0 & 0xFFFF) | (run1 << 16)
                run0 += c & 0xFFFF
                run1 += c >> 16
            run += tot0 + tot1
        nvalid = run
        assert nvalid == min(TILE, n - tile_base)
        # ---- staging: position inside the tile = warp's running offset of the digit + rank
        staged = [None] * nvalid
        for (warp, k, lane), pos in item.items():
            key = keys[tile_base + warp * 32 * IPT + k * 32 + lane]
            d = digit(key)
            off = (int(cnt2[warp, d >> 1]) >> (16 * (d & 1))) & 0xFFFF
            p = off + pos
            assert staged[p] is None
            staged[p] = key
        # ---- global bases and write-out
        gbase = bucket_start + prev_tiles - lstart
        for p in range(nvalid):
            dst = int(gbase[digit(staged[p])]) + p
            assert out[dst] is None
            out[dst] = staged[p]
        prev_tiles += tot
    return np.array(out, dtype=np.int64)


def check(keys, shift):
    got = pass9(keys, shift)
    d = (keys >> shift) & (RADIX - 1)
    want = keys[np.argsort(d, kind="stable")]
    assert np.array_equal(got, want)


def main():
    rng = np.random.default_rng(9)
    for n, hi in [(1, 1 << 27), (31, 1 << 27), (4096, 1 << 27), (4097, 1 << 9), (9000, 1 << 18), (12288, 3), (10000, 1 << 27)]:
        base = rng.integers(0, hi, n).astype(np.int64)
        keys = (base << 5) | rng.integers(0, 32, n)   # the low 5 bits tell entries with equal digits apart (stability)
        for shift in (5, 14, 23):
            check(keys, shift)
    print("k_radix_pass9 tile arithmetic: stable by digit in every case")


if __name__ == "__main__":
    main()
