// reduce_warp.cuh -- the duplicate-reduce of consolidate (reduce_by_key.cuh: k_reduce_by_key<MODE_CONSOLIDATE>) with a WARP
// per tile instead of a block per tile.
//
// Same contract: equal keys are adjacent and in insertion order after the stable sort, the head of every run folds its
// followers left to right -- acc = v0; acc += v1; ... (reference slib/spsparse/algorithm.hpp:277-313), so the sums are
// bit-identical to the reference's -- outputs are compacted through a decoupled look-back, the compressed row starts of
// the output (dim_beginnings, algorithm.hpp:74-118) come out as a by-product, runs of more than RK_LONG_RUN duplicates are
// handed to k_long_runs.
//
// Why: the block-per-tile kernel is bound by its own phases, not by HBM (ncu, profiles/r02_ncu_full_consolidate_kernels.csv:
// 37 % DRAM, issue slots 36 % busy, 10.8 barrier-stall cycles per issue): eight warps wait at a barrier for the slowest
// warp's loads, then for warp 0's look-back, then for the staging.  Here a warp owns 256 consecutive entries from load to
// store: four 16-byte loads per array and thread (a thread's 8 consecutive entries), heads and folds in registers, one warp
// scan, the warp's own look-back, a warp-private staging area for coalesced stores -- and no block barrier after the ticket,
// so a warp that waits (for memory, for its predecessors) never holds seven others.
#pragma once
#include "reduce_by_key.cuh"

constexpr int RW_IPT = 8;
constexpr int RW_TILE = 32 * RW_IPT;          // entries per warp
constexpr int RW_WARPS = 8;                   // warps per block; they share nothing but the ticket
constexpr int RW_THREADS = 32 * RW_WARPS;
constexpr int RW_SLOTS = RW_TILE + (RW_TILE >> 3) + 1;   // staging slots, one pad every 8 (rk_phys)
#ifndef RW_MIN_BLOCKS
#define RW_MIN_BLOCKS 4
#endif

// LOOKBACK = true: the tile's place in the output comes from a decoupled look-back over a.state (zeroed status words).
// LOOKBACK = false: a.state[tile] already holds the exclusive prefix (entries | rows << 31) -- the in-row sort counted the run
// heads and row heads of every tile of its output (radix_sort.cuh: sg_count_heads) and the counts were scanned; no warp ever
// waits for another.
template <bool LOOKBACK>
__global__ void __launch_bounds__(RW_THREADS, RW_MIN_BLOCKS) k_reduce_warp(ReduceArgs a) {
    __shared__ u64 s_keys[RW_WARPS][RW_SLOTS];
    __shared__ double s_vals[RW_WARPS][RW_SLOTS];
    __shared__ u32 s_ticket;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u32 tile = blockIdx.x * RW_WARPS + warp;
    if (LOOKBACK) {
        // tiles are numbered in the order the blocks start, so every predecessor of a tile is running or done (the look-back
        // cannot wait for a tile that has not been scheduled)
        if (tid == 0) s_ticket = atomicAdd(a.ticket, 1u);
        __syncthreads();
        tile = s_ticket * RW_WARPS + warp;
    }
    const u32 n = *a.n_ptr;
    const u64 base = (u64)tile * RW_TILE;
    if (base >= n) return;
    const u32 tile_n = (n - base < (u64)RW_TILE) ? (u32)(n - base) : (u32)RW_TILE;
    const bool want_rows = a.row_start != nullptr;
    const u32 first = lane * RW_IPT;  // tile-local index of my first entry
    const u32 mine = first < tile_n ? (tile_n - first < (u32)RW_IPT ? tile_n - first : (u32)RW_IPT) : 0;

    // ---- my 8 consecutive entries ------------------------------------------------------------------------------------------
    u64 key[RW_IPT];
    double acc[RW_IPT];
    if (tile_n == (u32)RW_TILE && ((((uintptr_t)a.keys) | ((uintptr_t)a.vals)) & 15u) == 0) {
        const ulonglong2 *kp = reinterpret_cast<const ulonglong2 *>(a.keys + base + first);
        const double2 *vp = reinterpret_cast<const double2 *>(a.vals + base + first);
#pragma unroll
        for (int j = 0; j < RW_IPT / 2; ++j) {
            const ulonglong2 t = __ldg(kp + j);
            key[2 * j] = t.x; key[2 * j + 1] = t.y;
        }
#pragma unroll
        for (int j = 0; j < RW_IPT / 2; ++j) {
            const double2 t = __ldg(vp + j);
            acc[2 * j] = t.x; acc[2 * j + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < RW_IPT; ++j) {
            key[j] = 0; acc[j] = 0.0;
            if ((u32)j < mine) { key[j] = a.keys[base + first + j]; acc[j] = a.vals[base + first + j]; }
        }
    }
    // the key before my first entry, the key after my last one
    u64 prev = __shfl_up_sync(SPB_FULL_MASK, key[RW_IPT - 1], 1);
    if (lane == 0) prev = base ? a.keys[base - 1] : 0;
    u64 next = __shfl_down_sync(SPB_FULL_MASK, key[0], 1);
    const bool more = base + first + RW_IPT < n;   // an entry follows my block
    if (lane == 31 && more) next = a.keys[base + RW_TILE];

    u32 head_bits = 0, rhead_bits = 0, defer_bits = 0;
    if (mine) {
        int cur = -1;  // my open run (index of its head among my entries), -1: none yet
#pragma unroll
        for (int j = 0; j < RW_IPT; ++j) {
            if ((u32)j < mine) {
                const u64 before = j ? key[j - 1] : prev;
                const bool head = (a.policy == POLICY_KEEP_ALL) || (base + first + j == 0) || (key[j] != before);
                if (head) {
                    head_bits |= 1u << j;
                    if (want_rows && ((base + first + j == 0) || (key[j] >> a.bits_lo) != (before >> a.bits_lo)))
                        rhead_bits |= 1u << j;
                    cur = j;
                } else if (cur >= 0) {  // follower of a run that started in my block: fold left to right
                    const double v = acc[j];
#pragma unroll
                    for (int h = 0; h < RW_IPT; ++h)
                        if (h == cur) {
                            if (a.policy == POLICY_ADD) acc[h] = __dadd_rn(acc[h], v);
                            else if (a.policy == POLICY_REPLACE) acc[h] = v;
                        }
                }
            }
        }
        // my last run continues past my block (seen from the neighbour's first key, no load): keep folding from memory --
        // the entries are in L1/L2, the neighbours have just loaded them
        if (cur >= 0 && mine == (u32)RW_IPT && more && next == key[RW_IPT - 1] &&
            (a.policy == POLICY_ADD || a.policy == POLICY_REPLACE)) {
            const u64 k0 = key[RW_IPT - 1];
            double sum = 0.0;
#pragma unroll
            for (int h = 0; h < RW_IPT; ++h) if (h == cur) sum = acc[h];
            u64 q = base + first + RW_IPT;  // global index of the next entry
            u32 steps = 0;
            while (q < n) {
                if (a.keys[q] != k0) break;
                const double vq = a.vals[q];
                if (a.policy == POLICY_ADD) sum = __dadd_rn(sum, vq); else sum = vq;
                ++q;
                if (++steps >= RK_LONG_RUN) { defer_bits |= 1u << cur; break; }
            }
#pragma unroll
            for (int h = 0; h < RW_IPT; ++h) if (h == cur) acc[h] = sum;
        }
    }
    // ---- slots inside the tile: one warp scan; the tile's place in the output: the warp's own look-back -------------------------
    const u32 emit_bits = head_bits;
    const u64 my = (u64)__popc(emit_bits) | ((u64)__popc(rhead_bits) << 32);
    const u64 incl = warp_incl_scan(my);
    const u64 total = __shfl_sync(SPB_FULL_MASK, incl, 31);
    const u64 before_me = incl - my;
    const u32 tile_out = (u32)(total & 0xffffffffull);
    // look-back value: entries in bits [0,31), rows in bits [31,62)
    const u64 packed = (total & 0xffffffffull) | ((total >> 32) << 31);
    const u64 excl0 = LOOKBACK ? lookback_exclusive(a.state, tile, packed) : a.state[tile];
    if (lane == 0 && base + RW_TILE >= n) {
        const u64 fin = excl0 + packed;
        const u32 n_out = (u32)(fin & 0x7fffffffull), n_rows = (u32)(fin >> 31);
        *a.out_count = n_out;
        if (want_rows) { *a.row_count = n_rows; a.row_start[n_rows] = n_out; }
    }
    const u64 excl_rows = excl0 >> 31;
    const u64 excl = excl0 & 0x7fffffffull;
    // ---- outputs: staged at their tile-local slots (warp-private), row starts and deferred runs, coalesced copy-out -------------
    u64 *const sk = s_keys[warp];
    double *const sv = s_vals[warp];
    {
        u32 slot = (u32)(before_me & 0xffffffffull);
        u64 rslot = excl_rows + (before_me >> 32);
#pragma unroll
        for (int j = 0; j < RW_IPT; ++j) {
            if ((emit_bits >> j) & 1u) {
                sk[rk_phys(slot)] = key[j];
                sv[rk_phys(slot)] = acc[j];
                if ((rhead_bits >> j) & 1u) {
                    a.row_start[rslot] = (u32)(excl + slot);
                    a.row_id[rslot] = (i32)(key[j] >> a.bits_lo);
                    ++rslot;
                }
                if ((defer_bits >> j) & 1u) {
                    const u32 t = atomicAdd(a.long_count, 1u);
                    if (t < a.long_cap) { a.long_list[2 * t] = (u32)(excl + slot); a.long_list[2 * t + 1] = (u32)(base + first + j); }
                }
                ++slot;
            }
        }
    }
    __syncwarp();
    const u64 lo_mask = (1ull << a.bits_lo) - 1;
    for (u32 t = lane; t < tile_out; t += 32) {
        const u64 k = sk[rk_phys(t)];
        a.out_hi[excl + t] = (i32)(k >> a.bits_lo);
        if (a.out_lo) a.out_lo[excl + t] = (i32)(k & lo_mask);
        a.out_val[excl + t] = sv[rk_phys(t)];
    }
}
