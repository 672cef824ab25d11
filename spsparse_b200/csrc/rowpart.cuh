// rowpart.cuh -- device side of the row-partitioned multiply on several GPUs of one node (one process per GPU).
//
// The reference is single-process; what is split here is its A-row loop (slib/spsparse/multiply_sparse.hpp:192-246 carries no
// state from one row of A to the next): rank r owns rows [row_lo[r], row_lo[r+1]) of op(A) and the same rows of B (B's rows
// are the inner index).  Every step each rank consolidates its shard of B, PUBLISHES it in compressed form (local row
// pointers + column + value of every entry, 12 B per entry) in a buffer that the other ranks have mapped (CUDA IPC over
// NVLink), and FETCHES the rows of B that its block of A can reference -- the interval hull [lo, hi] of the inner indices of
// its A entries: everything for a general matrix, its own shard plus a halo for a banded one -- with plain loads from the
// peers' memory, while consolidate(A) runs.  No host round trip is involved: sizes, offsets and the hand-shake (ready /
// done step counters written into the peers' memory) all stay on the devices.
#pragma once
#include "common.cuh"

constexpr int RP_MAX_RANKS = 16;
constexpr int RP_FLAG_WORDS = 2 * RP_MAX_RANKS;    // ready[RP_MAX_RANKS], done[RP_MAX_RANKS]
constexpr u64 RP_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;  // a peer that does not show up within 20 s: give up (error flag), never hang

// One rank's exported region as it is mapped in this process.  cols / vals point at the SHARD, which sits in the middle of a
// larger allocation: `slack` entries of room on either side take the halo rows fetched from the neighbouring ranks, so that a
// rank whose block of A only reaches a little beyond its own rows multiplies against [lower halo | own shard | upper halo]
// in place -- its own shard, consolidated straight into this buffer, is never copied.
struct RpRegion {
    u64 *flags;    // [0 .. RP_MAX_RANKS) ready[q]: last step whose shard rank q has published (q writes it into MY region)
                   // [RP_MAX_RANKS .. )  done[q]: last step in which rank q has finished reading MY shard
    u32 *ptr;      // [cap_rows + 1] entry offset, inside the shard, of the first entry of each of its rows (+ sentinel)
    i32 *cols;     // [cap_entries] (+ slack before and after)
    double *vals;  // [cap_entries] (+ slack before and after)
};

struct RpArgs {
    int rank, n_ranks;
    u64 step;                         // 1, 2, ...
    u64 row_lo[RP_MAX_RANKS + 1];
    RpRegion reg[RP_MAX_RANKS];       // reg[rank] is this rank's own region
    u32 slack;                        // entries of room before and after every shard
    u32 *error;                       // set to 1 when a wait timed out
};

__device__ __forceinline__ u64 rp_now_ns() {
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ u64 rp_ld_acquire_sys(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void rp_st_release_sys(u64 *p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spins until *p >= want; false (and *error = 1) on time-out
__device__ __forceinline__ bool rp_wait_ge(const u64 *p, u64 want, u32 *error) {
    const u64 t0 = rp_now_ns();
    while (rp_ld_acquire_sys(p) < want) {
        __nanosleep(200);
        if (rp_now_ns() - t0 > RP_TIMEOUT_NS) { *error = 1u; return false; }
    }
    return true;
}

// ---- publish: wait until every peer has finished reading this buffer's previous publication (two steps ago), copy, raise the
//      ready counters ------------------------------------------------------------------------------------------------------------
__global__ void k_rp_wait_done(RpArgs a) {
    const int q = threadIdx.x;
    // (two shard buffers alternate: this step's buffer was last read in step - 2)
    if (q < a.n_ranks && q != a.rank && a.step > 2)
        rp_wait_ge(a.reg[a.rank].flags + RP_MAX_RANKS + q, a.step - 2, a.error);
}

// local_ptr: rows_local + 1 offsets into (cols, vals) of the consolidated shard (spb_coo_dense_ptr_range); published re-based to 0
__global__ void __launch_bounds__(512) k_rp_publish(RpArgs a, const u32 *__restrict__ local_ptr, u32 rows_local,
                                                    const i32 *__restrict__ cols, const double *__restrict__ vals) {
    const RpRegion me = a.reg[a.rank];
    const u32 e0 = local_ptr[0], n = local_ptr[rows_local] - e0;
    const u64 stride = (u64)gridDim.x * blockDim.x, t0 = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    for (u64 t = t0; t <= rows_local; t += stride) me.ptr[t] = local_ptr[t] - e0;
    for (u64 t = t0; t < n; t += stride) {
        me.cols[t] = ld_stream_i32(cols + e0 + t);
        me.vals[t] = ld_stream_f64(vals + e0 + t);
    }
}

// the shard was consolidated straight into the region: only its row pointers are published.  Entries of rows outside the
// shard's range (local_ptr[0] != 0) would shift everything: the caller's partition is wrong, error 4.
__global__ void __launch_bounds__(512) k_rp_publish_ptr(RpArgs a, const u32 *__restrict__ local_ptr, u32 rows_local, u32 n_shard) {
    const RpRegion me = a.reg[a.rank];
    if (blockIdx.x == 0 && threadIdx.x == 0 && (local_ptr[0] != 0 || local_ptr[rows_local] != n_shard)) *a.error = 4u;
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t <= rows_local; t += (u64)gridDim.x * blockDim.x) me.ptr[t] = local_ptr[t];
}

__global__ void k_rp_signal_ready(RpArgs a) {
    const int q = threadIdx.x;
    __threadfence_system();
    if (q < a.n_ranks) rp_st_release_sys(a.reg[q].flags + a.rank, a.step);   // ready[rank] in rank q's region (also my own)
}

// ---- interval hull of the inner indices of this rank's block of A (raw, unconsolidated entries) ------------------------------
// hull[0] = min, hull[1] = max (initialised to ~0 / 0 by the host); an empty block leaves lo > hi = nothing to fetch
__global__ void k_rp_hull(const i32 *__restrict__ a_inner, u64 n, u64 *hull) {
    u32 lo = 0xffffffffu, hi = 0;
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (u64)gridDim.x * blockDim.x) {
        const u32 j = (u32)ld_stream_i32(a_inner + t);
        lo = min(lo, j);
        hi = max(hi, j);
    }
    lo = __reduce_min_sync(SPB_FULL_MASK, lo);
    hi = __reduce_max_sync(SPB_FULL_MASK, hi);
    if (lane_id() == 0 && lo <= hi) {
        atomicMin((ull *)&hull[0], (ull)lo);
        atomicMax((ull *)&hull[1], (ull)hi);
    }
}
__global__ void k_rp_hull_set(u64 *hull, u64 lo, u64 hi) { hull[0] = lo; hull[1] = hi; }

// ---- fetch: the rows lo..hi of B from whoever owns them, concatenated in row order ----------------------------------------------
// g_ptr is indexed by the ABSOLUTE row: g_ptr[j] for j in [lo, hi + 1] is written (offsets into g_cols / g_vals), nothing else --
// the multiply only ever looks up the inner indices of this rank's A entries, which lie inside the hull.  Work and traffic are
// O(rows and entries fetched); only the allocation is as long as B has rows.
// info[0] = entries fetched, info[1] = rows fetched, info[2] = blocks finished (zeroed by the host)
constexpr int RP_PULL_THREADS = 512;
constexpr int RP_PULL_UNROLL = 8;
// info[0] = entries fetched, info[1] = rows fetched, info[2] = blocks finished (zeroed by the host), info[3] = outcome:
// 1 = done in place, 2 = done by copying, 3 = the halos do not fit the slack (nothing done, the peers NOT released: the host
// launches the copying variant)
// Waits, with ONE warp, for the ready counters the fetch below will need (the peers whose rows lie inside the hull, and the own
// shard).  The fetch kernel waits for them too -- every block works its plan out for itself -- but its blocks are 512 threads with
// eight 16-byte loads in flight each: 32 of them parked on a late neighbour took half the registers of 32 SMs away from the
// consolidate(A) that runs meanwhile (measured at N = 8: +0.3 ms on the ranks that waited).  Launched in front of it on the side
// stream, this kernel does the waiting; the fetch then finds the counters raised.
__global__ void __launch_bounds__(32) k_rp_wait_ready(RpArgs a, const u64 *__restrict__ hull) {
    const int g = (int)threadIdx.x;
    const u64 lo = hull[0], hi = hull[1];
    if (g < a.n_ranks) {
        bool need = g == a.rank;
        if (lo <= hi) need = need || max(lo, a.row_lo[g]) < min(hi + 1, a.row_lo[g + 1]);
        if (need) rp_wait_ge(a.reg[a.rank].flags + g, a.step, a.error);
    }
}

template <bool IN_PLACE>
__global__ void __launch_bounds__(RP_PULL_THREADS) k_rp_pull(RpArgs a, const u64 *__restrict__ hull, u32 *g_ptr, i32 *g_cols,
                                                             double *g_vals, u64 cap_entries, u64 *info) {
    __shared__ u64 s_rlo[RP_MAX_RANKS], s_rhi[RP_MAX_RANKS];   // rows [rlo, rhi) of peer g are fetched
    __shared__ u32 s_elo[RP_MAX_RANKS], s_cnt[RP_MAX_RANKS];   // its entries [elo, elo + cnt)
    __shared__ u64 s_base[RP_MAX_RANKS + 1];                   // where they go
    __shared__ u32 s_last, s_fit;
    const u32 tid = threadIdx.x;
    const u64 lo = hull[0], hi = hull[1];
    if (tid < 32) {
        // every block works the plan out for itself (two pointer values per peer): no block waits for another block
        const int g = (int)tid;
        u64 rlo = 0, rhi = 0;
        u32 elo = 0, cnt = 0, n_own = 0;
        if (g < a.n_ranks && lo <= hi) {
            rlo = max(lo, a.row_lo[g]);
            rhi = min(hi + 1, a.row_lo[g + 1]);
            if (rlo < rhi) {
                bool ok = rp_wait_ge(a.reg[a.rank].flags + g, a.step, a.error);   // peer g has published this step's shard
                if (ok) {
                    const u32 *p = a.reg[g].ptr;
                    elo = __ldcg(p + (rlo - a.row_lo[g]));
                    cnt = __ldcg(p + (rhi - a.row_lo[g])) - elo;
                } else rhi = rlo;
            } else rhi = rlo;
        }
        if (IN_PLACE) {
            // own shard stays where it is (offset 0 of my region's shard); lower halos end right before it, upper ones start
            // right after its last entry
            if (g == a.rank) {
                rp_wait_ge(a.reg[a.rank].flags + g, a.step, a.error);
                n_own = __ldcg(a.reg[g].ptr + (a.row_lo[g + 1] - a.row_lo[g]));
            }
            n_own = __shfl_sync(SPB_FULL_MASK, n_own, a.rank);
            const u32 c_lo = g < a.rank ? cnt : 0, c_hi = (g > a.rank && g < a.n_ranks) ? cnt : 0;
            const u64 in_lo = warp_incl_scan((u64)c_lo), in_hi = warp_incl_scan((u64)c_hi);
            const u64 tot_lo = __shfl_sync(SPB_FULL_MASK, in_lo, 31), tot_hi = __shfl_sync(SPB_FULL_MASK, in_hi, 31);
            // positions relative to the first entry of my shard (negative = inside the lower slack), biased by slack
            u64 base = a.slack;                                               // own
            if (g < a.rank) base = a.slack - tot_lo + (in_lo - c_lo);
            else if (g > a.rank) base = (u64)a.slack + n_own + (in_hi - c_hi);
            if (g < a.n_ranks) { s_rlo[g] = rlo; s_rhi[g] = rhi; s_elo[g] = elo; s_cnt[g] = cnt; s_base[g] = base; }
            if (g == 0) { s_base[a.n_ranks] = tot_lo + tot_hi; s_fit = (tot_lo <= a.slack && tot_hi <= a.slack) ? 1u : 0u; }
        } else {
            const u64 incl = warp_incl_scan((u64)cnt);
            if (g < a.n_ranks) { s_rlo[g] = rlo; s_rhi[g] = rhi; s_elo[g] = elo; s_cnt[g] = cnt; s_base[g] = incl - cnt; }
            if (g == a.n_ranks - 1) s_base[a.n_ranks] = incl;
            if (g == 0) s_fit = 1u;
        }
    }
    __syncthreads();
    const u64 total = s_base[a.n_ranks];     // entries copied
    if (IN_PLACE && !s_fit) {                // the host switches to the copying variant; the peers stay un-released until then
        if (blockIdx.x == 0 && tid == 0) info[3] = 3;
        return;
    }
    const bool overflow = !IN_PLACE && total > cap_entries;   // never overrun the buffers: nothing is copied, the host reports it
    if (overflow && tid == 0) *a.error = 2u;
    // IN_PLACE: everything is addressed relative to (my shard - slack) inside my own region
    i32 *const out_c = IN_PLACE ? a.reg[a.rank].cols - a.slack : g_cols;
    double *const out_v = IN_PLACE ? a.reg[a.rank].vals - a.slack : g_vals;
    // entries: chunks of RP_PULL_THREADS * RP_PULL_UNROLL, all loads of a chunk in flight before the first store
    constexpr u64 CHUNK = (u64)RP_PULL_THREADS * RP_PULL_UNROLL;
    for (int g = 0; g < a.n_ranks && !overflow; ++g) {
        const u32 cnt = s_cnt[g];
        const u64 rlo = s_rlo[g], rhi = s_rhi[g], r0 = a.row_lo[g];
        if (rlo >= rhi) continue;
        const bool own_in_place = IN_PLACE && g == a.rank;
        if (cnt && !own_in_place) {
            const i32 *sc = a.reg[g].cols + s_elo[g];
            const double *sv = a.reg[g].vals + s_elo[g];
            i32 *dc = out_c + s_base[g];
            double *dv = out_v + s_base[g];
            // 16-byte loads from the peer (NVLink moves 512 B per warp request instead of 128 / 256): the source is aligned to
            // 16 bytes from entry `head` on (the shard starts on a 256-byte boundary), the destination is written entry by entry
            const u32 head_c = min(cnt, (4u - (s_elo[g] & 3u)) & 3u), quads = (cnt - head_c) / 4, tail_c = head_c + quads * 4;
            const u32 head_v = min(cnt, s_elo[g] & 1u), pairs = (cnt - head_v) / 2, tail_v = head_v + pairs * 2;
            const int4 *sc4 = reinterpret_cast<const int4 *>(sc + head_c);
            const double2 *sv2 = reinterpret_cast<const double2 *>(sv + head_v);
            for (u64 q0 = (u64)blockIdx.x * CHUNK; q0 < quads; q0 += (u64)gridDim.x * CHUNK) {
                int4 k[RP_PULL_UNROLL];
#pragma unroll
                for (int u = 0; u < RP_PULL_UNROLL; ++u) {
                    const u64 q = q0 + (u64)u * RP_PULL_THREADS + tid;
                    if (q < quads) k[u] = __ldcg(sc4 + q);   // L2 only: the owner rewrites these every other step
                }
#pragma unroll
                for (int u = 0; u < RP_PULL_UNROLL; ++u) {
                    const u64 q = q0 + (u64)u * RP_PULL_THREADS + tid;
                    if (q < quads) { i32 *d = dc + head_c + 4 * q; d[0] = k[u].x; d[1] = k[u].y; d[2] = k[u].z; d[3] = k[u].w; }
                }
            }
            for (u64 q0 = (u64)blockIdx.x * CHUNK; q0 < pairs; q0 += (u64)gridDim.x * CHUNK) {
                double2 v[RP_PULL_UNROLL];
#pragma unroll
                for (int u = 0; u < RP_PULL_UNROLL; ++u) {
                    const u64 q = q0 + (u64)u * RP_PULL_THREADS + tid;
                    if (q < pairs) v[u] = __ldcg(sv2 + q);
                }
#pragma unroll
                for (int u = 0; u < RP_PULL_UNROLL; ++u) {
                    const u64 q = q0 + (u64)u * RP_PULL_THREADS + tid;
                    if (q < pairs) { double *d = dv + head_v + 2 * q; d[0] = v[u].x; d[1] = v[u].y; }
                }
            }
            if (blockIdx.x == 0) {   // the few entries in front of and behind the aligned body
                if (tid < head_c) dc[tid] = __ldcg(sc + tid);
                if (tid < cnt - tail_c) dc[tail_c + tid] = __ldcg(sc + tail_c + tid);
                if (tid < head_v) dv[tid] = __ldcg(sv + tid);
                if (tid < cnt - tail_v) dv[tail_v + tid] = __ldcg(sv + tail_v + tid);
            }
        }
        // row pointers of the fetched rows, re-based to where the entries are (own rows in place: offset slack)
        const u32 *p = a.reg[g].ptr;
        const u32 shift = own_in_place ? a.slack : (u32)s_base[g] - s_elo[g];   // wraps consistently in 32 bits
        for (u64 j = rlo + (u64)blockIdx.x * blockDim.x + tid; j < rhi; j += (u64)gridDim.x * blockDim.x)
            g_ptr[j] = __ldcg(p + (j - r0)) + shift;
    }
    if (blockIdx.x == 0 && tid == 0 && lo <= hi && !overflow) {
        // end of the last fetched row
        int gl = 0;
        for (int g = 0; g < a.n_ranks; ++g) if (s_rlo[g] < s_rhi[g]) gl = g;
        const u32 shift = (IN_PLACE && gl == a.rank) ? a.slack : (u32)s_base[gl] - s_elo[gl];
        g_ptr[hi + 1] = s_elo[gl] + s_cnt[gl] + shift;
    }
    // the last block to finish tells every peer that this rank is done with its shard for this step
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd((ull *)&info[2], 1ull) == (ull)gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        if (tid == 0) {
            u64 fetched = 0;
            for (int g = 0; g < a.n_ranks; ++g) fetched += s_cnt[g];
            info[0] = fetched; info[1] = lo <= hi ? hi - lo + 1 : 0; info[3] = IN_PLACE ? 1 : 2;
        }
        __threadfence_system();
        if (tid < (u32)a.n_ranks && (int)tid != a.rank) rp_st_release_sys(a.reg[tid].flags + RP_MAX_RANKS + a.rank, a.step);
    }
}

// ---- scale vector for the fetched range only: dense[j] / mask[j] for j in [lo, hi] -----------------------------------------
__global__ void k_rp_zero_range(const u64 *__restrict__ hull, double *dense, unsigned char *mask) {
    const u64 lo = hull[0], hi = hull[1];
    if (lo > hi) return;
    for (u64 j = lo + (u64)blockIdx.x * blockDim.x + threadIdx.x; j <= hi; j += (u64)gridDim.x * blockDim.x) {
        dense[j] = 0.0;
        mask[j] = 0;
    }
}
__global__ void k_rp_densify_range(const i32 *__restrict__ idx, const double *__restrict__ val, u64 n, const u64 *__restrict__ hull,
                                   double *dense, unsigned char *mask, u32 *bad) {
    const u64 lo = hull[0], hi = hull[1];
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (u64)gridDim.x * blockDim.x) {
        const u64 j = (u64)(u32)idx[t];
        if (t && idx[t - 1] >= idx[t]) *bad = 1u;
        if (j >= lo && j <= hi) { dense[j] = val[t]; mask[j] = 1; }
    }
}
