"""CPU suite: the per-tile arithmetic of k_radix_pass9 (9-bit digits, 16-bit warp counters packed two to a word, one owner
thread per digit pair; spsparse_b200/csrc/radix_sort9.cuh) emulated in Python: the pass must be the stable sort of the keys
by the digit (tools/emulate_radix_pass9.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import emulate_radix_pass9 as emu  # noqa: E402


def test_pass_is_stable_by_digit():
    rng = np.random.default_rng(99)
    for n, hi in [(1, 1 << 27), (33, 1 << 27), (4096, 1 << 27), (4097, 1 << 9), (8200, 1 << 18), (8192, 2)]:
        keys = (rng.integers(0, hi, n).astype(np.int64) << 5) | rng.integers(0, 32, n)
        for shift in (5, 14, 23):
            emu.check(keys, shift)


def test_three_passes_sort_a_27_bit_key():
    rng = np.random.default_rng(27)
    keys = rng.integers(0, 1 << 27, 9000).astype(np.int64)
    tagged = (keys << 20) | np.arange(9000)   # the tag tells equal keys apart: the three passes must keep their order
    out = tagged
    for p in range(3):
        out = emu.pass9(out, 20 + 9 * p)
    assert np.array_equal(out, tagged[np.argsort(keys, kind="stable")])
