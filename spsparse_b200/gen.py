"""numpy copies of the device generators in csrc/gen.cuh, for tests and for
feeding the CPU baseline the same inputs.  splitmix64 on a counter, uint64 wrap-around."""
from __future__ import annotations

import numpy as np

SCRAMBLE = 2654435761
_U = np.uint64


def mix64(x):
    with np.errstate(over="ignore"):
        z = np.asarray(x, dtype=_U) + _U(0x9E3779B97F4A7C15)
        z = (z ^ (z >> _U(30))) * _U(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U(27))) * _U(0x94D049BB133111EB)
        return z ^ (z >> _U(31))


def u01(x):
    return (mix64(x) >> _U(11)).astype(np.float64) * 2.0 ** -53


def _scramble_big(e, n):
    """(e * SCRAMBLE) mod n in uint64 without overflow (e, n < 2^33): split the multiplier."""
    e = np.asarray(e, dtype=_U)
    n = _U(n)
    s1, s0 = _U(SCRAMBLE >> 16), _U(SCRAMBLE & 0xFFFF)
    hi = (e * s1) % n
    return ((hi * _U(1 << 16) + e * s0) % n).astype(np.int64)


def dup_coo(seed, i0, n, ubase, bits, zero_every=0):
    with np.errstate(over="ignore"):
        i = np.arange(i0, i0 + n, dtype=_U)
        s = np.where(i < _U(ubase), i, mix64((_U(seed) ^ _U(0xD0B1E)) + i) % _U(ubase))
        mask = _U((1 << bits) - 1)
        row = (mix64(_U(seed) + _U(2) * s) & mask).astype(np.int32)
        col = (mix64(_U(seed) + _U(2) * s + _U(1)) & mask).astype(np.int32)
        val = 0.5 + u01((_U(seed) ^ _U(0xA11CE)) + i)
        if zero_every:
            val[mix64((_U(seed) ^ _U(0x2E80)) + i) % _U(zero_every) == 0] = 0.0
    return (1 << bits, 1 << bits), [row, col], val


def banded(seed, m, r0, r1):
    n = (r1 - r0) * 5
    with np.errstate(over="ignore"):
        e = np.arange(n, dtype=np.int64)
        i = r0 + e // 5
        d = e % 5 - 2
        c = i + d
        val = 0.5 + u01(_U(seed) + (5 * i + d + 2).astype(_U))
        val[(c < 0) | (c >= m)] = 0.0
        c = np.clip(c, 0, m - 1)
    slot = _scramble_big(e, n)
    row = np.empty(n, np.int32); col = np.empty(n, np.int32); v = np.empty(n, np.float64)
    row[slot], col[slot], v[slot] = i, c, val
    return (m, m), [row, col], v


def regrid(seed, ny, nx, gy, gx):
    n = ny * nx * 4
    with np.errstate(over="ignore"):
        e = np.arange(n, dtype=np.int64)
        r, q = e // 4, e % 4
        y, x = r // nx, r % nx
        yy = np.minimum((y * gy) // ny + (q >> 1), gy - 1)
        xx = np.minimum((x * gx) // nx + (q & 1), gx - 1)
        val = 0.1 + 0.9 * u01(_U(seed) + e.astype(_U))
    slot = _scramble_big(e, n)
    row = np.empty(n, np.int32); col = np.empty(n, np.int32); v = np.empty(n, np.float64)
    row[slot], col[slot], v[slot] = r, yy * gx + xx, val
    return (ny * nx, gy * gx), [row, col], v


def rmat(seed, scale, nedges):
    with np.errstate(over="ignore"):
        e = np.arange(nedges, dtype=_U)
        r = np.zeros(nedges, np.int64); c = np.zeros(nedges, np.int64)
        for l in range(scale):
            u = u01(_U(seed) + _U(scale) * e + _U(l))
            r = (r << 1) | (u >= 0.76)
            c = (c << 1) | (((u >= 0.57) & (u < 0.76)) | (u >= 0.95))
        val = 0.5 + u01((_U(seed) ^ _U(0x4A77)) + e)
    return (1 << scale, 1 << scale), [r.astype(np.int32), c.astype(np.int32)], val


def vector(seed, dim):
    with np.errstate(over="ignore"):
        j = np.arange(dim, dtype=_U)
        return (dim,), [j.astype(np.int32)], 0.5 + u01(_U(seed) + j)
