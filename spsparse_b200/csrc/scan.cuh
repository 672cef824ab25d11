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
// predecessors 32 at a time, publishes the inclusive prefix and returns the exclusive prefix.
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
        i64 idx = top - (i64)lane;
        u64 w = (idx >= 0) ? ld_relaxed_u64(&state[idx]) : LB_FLAG_INCL;
        while (__any_sync(SPB_FULL_MASK, LB_FLAG(w) == 0)) {
            if (LB_FLAG(w) == 0) w = ld_relaxed_u64(&state[idx]);
        }
        u32 incl = __ballot_sync(SPB_FULL_MASK, LB_FLAG(w) == 2);
        u64 v = LB_VALUE(w);
        if (incl) {
            u32 first = __ffs(incl) - 1;
            if (lane > first) v = 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SPB_FULL_MASK, v, o);
        excl += v;
        if (incl) break;
        top -= 32;
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
