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


def pull_in_place(row_lo, ptrs, cols, vals, lo, hi, m, rank, slack):
    """The IN_PLACE variant: rank's own shard sits at offset `slack` of its buffer and is never copied; the rows fetched from
    lower ranks end right before it, those from higher ranks start right after its last entry.  Returns None when the halos
    do not fit the slack, else (g_ptr, buf_cols, buf_vals) with buf_* the buffer addressed from (shard - slack)."""
    n = len(ptrs)
    rlo, rhi, elo, cnt = [0] * n, [0] * n, [0] * n, [0] * n
    for g in range(n):
        if lo <= hi:
            rlo[g] = max(lo, row_lo[g]); rhi[g] = min(hi + 1, row_lo[g + 1])
            if rlo[g] < rhi[g]:
                elo[g] = int(ptrs[g][rlo[g] - row_lo[g]]); cnt[g] = int(ptrs[g][rhi[g] - row_lo[g]]) - elo[g]
            else:
                rhi[g] = rlo[g]
    n_own = int(ptrs[rank][-1])
    tot_lo, tot_hi = sum(cnt[:rank]), sum(cnt[rank + 1:])
    if tot_lo > slack or tot_hi > slack:
        return None
    base = [0] * n
    run = 0
    for g in range(rank):
        base[g] = slack - tot_lo + run; run += cnt[g]
    base[rank] = slack
    run = 0
    for g in range(rank + 1, n):
        base[g] = slack + n_own + run; run += cnt[g]
    buf_c = np.full(2 * slack + n_own + 8, -7, dtype=np.int64); buf_v = np.zeros(2 * slack + n_own + 8)
    buf_c[slack:slack + n_own] = cols[rank]; buf_v[slack:slack + n_own] = vals[rank]      # consolidated in place
    g_ptr = np.full(m + 2, -1, dtype=np.int64)
    gl = None
    for g in range(n):
        if rlo[g] >= rhi[g]:
            continue
        gl = g
        own = g == rank
        if cnt[g] and not own:
            buf_c[base[g]:base[g] + cnt[g]] = cols[g][elo[g]:elo[g] + cnt[g]]
            buf_v[base[g]:base[g] + cnt[g]] = vals[g][elo[g]:elo[g] + cnt[g]]
        shift = slack if own else (base[g] - elo[g]) & 0xFFFFFFFF
        for j in range(rlo[g], rhi[g]):
            g_ptr[j] = (int(ptrs[g][j - row_lo[g]]) + shift) & 0xFFFFFFFF
    if lo <= hi and gl is not None:
        shift = slack if gl == rank else (base[gl] - elo[gl]) & 0xFFFFFFFF
        g_ptr[hi + 1] = (elo[gl] + cnt[gl] + shift) & 0xFFFFFFFF
    return g_ptr, buf_c, buf_v
