"""CPU emulation of the plan arithmetic of k_rp_pull (spsparse_b200/csrc/rowpart.cuh): which rows / entries of which peer a
rank fetches for the hull [lo, hi], where they land, and the row pointer indexed by the ABSOLUTE row.  Same formulas as the
kernel, numpy in place of threads; tests/test_multi_rank_cpu.py checks it against a direct construction."""
import numpy as np


def pull(row_lo, ptrs, cols, vals, lo, hi, m):
    """row_lo[g]..row_lo[g+1]: rows of peer g; ptrs[g]: its local row pointers (rows_g + 1, starting at 0); cols / vals: its
    entries.  Returns (g_ptr [m + 2] with only [lo, hi + 1] written, others = -1; g_cols; g_vals; total)."""
    n = len(ptrs)
    rlo, rhi, elo, cnt = [0] * n, [0] * n, [0] * n, [0] * n
    for g in range(n):
        if lo <= hi:
            rlo[g] = max(lo, row_lo[g]); rhi[g] = min(hi + 1, row_lo[g + 1])
            if rlo[g] < rhi[g]:
                elo[g] = int(ptrs[g][rlo[g] - row_lo[g]]); cnt[g] = int(ptrs[g][rhi[g] - row_lo[g]]) - elo[g]
            else:
                rhi[g] = rlo[g]
    base = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    total = int(base[-1])
    g_ptr = np.full(m + 2, -1, dtype=np.int64)
    g_cols = np.empty(total, dtype=np.int32); g_vals = np.empty(total, dtype=np.float64)
    for g in range(n):
        if not cnt[g] and rlo[g] >= rhi[g]:
            continue
        g_cols[base[g]:base[g] + cnt[g]] = cols[g][elo[g]:elo[g] + cnt[g]]
        g_vals[base[g]:base[g] + cnt[g]] = vals[g][elo[g]:elo[g] + cnt[g]]
        shift = (int(base[g]) - elo[g]) & 0xFFFFFFFF                       # 32-bit wrap, as in the kernel
        for j in range(rlo[g], rhi[g]):
            g_ptr[j] = (int(ptrs[g][j - row_lo[g]]) + shift) & 0xFFFFFFFF
    if lo <= hi:
        g_ptr[hi + 1] = total
    return g_ptr, g_cols, g_vals, total
