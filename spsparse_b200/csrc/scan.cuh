// scan.cuh -- single-pass device-wide exclusive scan (decoupled look-back), and the look-back
// helper shared with the duplicate-reduce / compaction kernels.
#pragma once
#include "common.cuh"

// Tile status word: [63:62] flag, [61:0] value.
#define LB_FLAG_AGG (1ull << 62)   // value = this tile's aggregate
#define LB_FLAG_INCL (2ull << 62)  // value = inclusive prefix up to and including this tile
#define LB_VALUE(w) ((w) & ((1ull << 62) - 1))
#define LB_FLAG(w) ((w) >> 62)

// Called by ALL 32 lanes of one warp.  Publishes `aggregate` for `tile`, walks back over the
// predecessors 32 * LB_WIDE at a time (every lane holds LB_WIDE consecutive status words, all loads of a round in
// flight together), publishes the inclusive prefix and returns the exclusive prefix.
// LB_WIDE = 1 (32 predecessors per round) is the measured optimum: the nearest inclusive prefix is almost always within
// the first 32 predecessors, and 128 / 256 per round made the reduce pass slower (5.61 -> 6.50 -> 7.63 ms on 5e8 entries,
// profiles/r02_notes.md) -- the barrier stalls of that kernel are not this walk.
#ifndef LB_WIDE
#define LB_WIDE 1
#endif
__device__ __forceinline__ u64 lookback_exclusive(u64 *state, u32 tile, u64 aggregate) {
    const u32 lane = lane_id();
    if (tile == 0) {
        if (lane == 0) st_relaxed_u64(&state[0], LB_FLAG_INCL | aggregate);
        return 0;
    }
    if (lane == 0) st_relaxed_u64(&state[tile], LB_FLAG_AGG | aggregate);
    u64 excl = 0;
    i64 top = (i64)tile - 1;
    for (;;) {
        // word u of lane l is the predecessor at distance l * LB_WIDE + u from `top`: nearer predecessors in lower lanes
        u64 w[LB_WIDE];
#pragma unroll
        for (int u = 0; u < LB_WIDE; ++u) {
            const i64 idx = top - (i64)(lane * LB_WIDE + u);
            w[u] = (idx >= 0) ? ld_relaxed_u64(&state[idx]) : LB_FLAG_INCL;
        }
        for (;;) {
            bool missing = false;
#pragma unroll
            for (int u = 0; u < LB_WIDE; ++u) missing |= LB_FLAG(w[u]) == 0;
            if (!__any_sync(SPB_FULL_MASK, missing)) break;
#pragma unroll
            for (int u = 0; u < LB_WIDE; ++u)
                if (LB_FLAG(w[u]) == 0) w[u] = ld_relaxed_u64(&state[top - (i64)(lane * LB_WIDE + u)]);
        }
        // my words up to and including my first inclusive one
        u64 v = 0;
        bool has_incl = false;
#pragma unroll
        for (int u = 0; u < LB_WIDE; ++u)
            if (!has_incl) {
                v += LB_VALUE(w[u]);
                has_incl = LB_FLAG(w[u]) == 2;
            }
        const u32 incl = __ballot_sync(SPB_FULL_MASK, has_incl);
        if (incl) {
            const u32 first = __ffs(incl) - 1;   // the nearest inclusive prefix is in this lane: farther lanes do not count
            if (lane > first) v = 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SPB_FULL_MASK, v, o);
        excl += v;
        if (incl) break;
        top -= 32 * LB_WIDE;
    }
    if (lane == 0) st_relaxed_u64(&state[tile], LB_FLAG_INCL | (excl + aggregate));
    return excl;
}

constexpr int SC_THREADS = 256;
constexpr int SC_IPT = 8;
constexpr int SC_TILE = SC_THREADS * SC_IPT;
constexpr int SC_WARPS = SC_THREADS / 32;

// out[i] = sum(in[0..i)) for i < n, and out[n] = total.  `state` needs ceil(n/SC_TILE) zeroed
// words, `ticket` one zeroed u32.
template <typename InT, typename OutT>
__global__ void __launch_bounds__(SC_THREADS) k_exclusive_scan(const InT *__restrict__ in,
                                                               OutT *__restrict__ out, u64 n,
                                                               u64 *state, u32 *ticket) {
    __shared__ u32 s_tile;
    __shared__ u64 s_part[SC_IPT * SC_WARPS];
    __shared__ u64 s_excl;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    const u64 base = (u64)tile * SC_TILE;
    if (base >= n) return;

    u64 v[SC_IPT], incl[SC_IPT];
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) {
        u64 i = base + (u64)k * SC_THREADS + tid;
        v[k] = (i < n) ? (u64)in[i] : 0;
    }
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) {
        incl[k] = warp_incl_scan(v[k]);
        if (lane == 31) s_part[k * SC_WARPS + warp] = incl[k];
    }
    __syncthreads();
    if (warp == 0) {
        // 64 partials, two per lane, in (k, warp) order
        u64 a = s_part[2 * lane], b = s_part[2 * lane + 1];
        u64 s = warp_incl_scan(a + b);
        u64 total = __shfl_sync(SPB_FULL_MASK, s, 31);
        s_part[2 * lane] = s - a - b;
        s_part[2 * lane + 1] = s - b;
        u64 excl = lookback_exclusive(state, tile, total);
        if (lane == 0) {
            s_excl = excl;
            if (base + SC_TILE >= n) out[n] = (OutT)(excl + total);
        }
    }
    __syncthreads();
    const u64 excl = s_excl;
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) {
        u64 i = base + (u64)k * SC_THREADS + tid;
        if (i < n) out[i] = (OutT)(excl + s_part[k * SC_WARPS + warp] + incl[k] - v[k]);
    }
}
