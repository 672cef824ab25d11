// reduce_by_key.cuh -- segmented duplicate-reduce + compaction over sorted packed keys.
//
// Replaces the merge loop of spsparse::consolidate (reference slib/spsparse/algorithm.hpp:277-313):
// equal keys are adjacent after the stable sort and still in insertion order, so the head of
// every run folds its followers left to right -- acc = v0; acc += v1; ... -- which is the
// reference's association order, hence bit-identical sums.  Runs longer than RK_LONG_RUN are
// finished by k_long_runs, one warp per run, in the same left-to-right order (bit-identical too).
// Output slots come from a single-pass decoupled look-back over the tiles' head counts.
//
// The same kernel is the "compress" step of expand-sort-compress in multiply (MODE_ESC): there
// the folded sum is the dot product of multiply_sparse.hpp:219-236, dropped when exactly zero
// (:238), scaled as ((sum*C)*a_scale)*b_scale (:242).
#pragma once
#include "scan.cuh"

#ifndef RK_THREADS_N
#define RK_THREADS_N 256   // threads per block of the reduce pass (same-box A/B builds: -DRK_THREADS_N=128 -DRK_MIN_BLOCKS=10)
#endif
constexpr int RK_THREADS = RK_THREADS_N;
constexpr int RK_IPT = 8;
constexpr int RK_TILE = RK_THREADS * RK_IPT;
constexpr int RK_WARPS = RK_THREADS / 32;
constexpr u32 RK_LONG_RUN = 256;

enum { POLICY_LEAVE_ALONE = 0, POLICY_ADD = 1, POLICY_REPLACE = 2, POLICY_KEEP_ALL = 3 };
enum { MODE_CONSOLIDATE = 0, MODE_ESC = 1 };

struct ReduceArgs {
    const u64 *keys;
    const double *vals;
    const u32 *n_ptr;
    int bits_lo;
    int policy;
    // output (MODE_CONSOLIDATE: hi/lo index vectors in sort-dimension order)
    i32 *out_hi;
    i32 *out_lo;  // nullptr for rank 1
    double *out_val;
    u32 *out_count;
    u64 *state;   // look-back words, zeroed
    u32 *ticket;  // zeroed
    // deferred long runs: pairs (output slot, first entry)
    u32 *long_list;
    u32 *long_count;
    u32 long_cap;
    // optional by-product (MODE_CONSOLIDATE, rank 2): compressed rows of the OUTPUT, i.e. dim_beginnings
    // (algorithm.hpp:74-118) for free while the keys are in shared memory
    u32 *row_start;   // [rows+1] offset of the first output entry of every non-empty leading index
    i32 *row_id;      // [rows]
    u32 *row_count;
    // MODE_ESC: key hi = compressed row number; emit (row number, k, scaled sum)
    const i32 *row_ids;  // compressed row -> row index i
    i32 row_base;        // key hi is relative to this compressed row number
    const double *si;    // dense a_scale per row index, or nullptr
    const double *sk;    // dense b_scale per column, or nullptr (0 => column excluded)
    double C;
};

// Shared-memory layout of a tile: logical slot L holds entry (tile_base + L - 1), so L = 0 is the
// predecessor of the tile and L = RK_TILE + 1 its successor.  One pad word every 8 slots makes the
// per-thread blocks of 8 consecutive entries (stride 9 words) conflict-free.
__device__ __forceinline__ u32 rk_phys(u32 L) { return L + (L >> 3); }
constexpr int RK_SLOTS = RK_TILE + 2 + ((RK_TILE + 2) >> 3) + 1;

#ifndef RK_MIN_BLOCKS
#define RK_MIN_BLOCKS 5
#endif
template <int MODE>
__global__ void __launch_bounds__(RK_THREADS, RK_MIN_BLOCKS) k_reduce_by_key(ReduceArgs a) {
    __shared__ u64 s_keys[RK_SLOTS];     // input keys; later: staged output keys
    __shared__ double s_vals[RK_SLOTS];  // input values; later: staged output values
    __shared__ u64 s_warp[RK_WARPS];     // per-warp totals: entries | rows << 32
    __shared__ u32 s_tile;
    __shared__ u64 s_excl;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool want_rows = (MODE == MODE_CONSOLIDATE) && (a.row_start != nullptr);
    if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    const u32 n = *a.n_ptr;
    const u64 base = (u64)tile * RK_TILE;
    if (base >= n) return;
    const u32 tile_n = (n - base < (u64)RK_TILE) ? (u32)(n - base) : (u32)RK_TILE;

    // ---- coalesced load of the tile (+ predecessor and successor keys) ----------------------------
#pragma unroll
    for (int k = 0; k < RK_IPT; ++k) {
        u32 p = (u32)k * RK_THREADS + tid;
        if (p < tile_n) {
            s_keys[rk_phys(p + 1)] = ld_stream_u64(a.keys + base + p);
            s_vals[rk_phys(p + 1)] = ld_stream_f64(a.vals + base + p);
        }
    }
    if (tid == 0) s_keys[0] = base ? a.keys[base - 1] : 0;
    if (tid == 32 && base + RK_TILE < n) s_keys[rk_phys(RK_TILE + 1)] = a.keys[base + RK_TILE];
    __syncthreads();

    // ---- each thread owns 8 consecutive entries -----------------------------------------------------
    const u32 first = tid * RK_IPT;  // tile-local index of my first entry
    const u32 mine = first < tile_n ? (tile_n - first < (u32)RK_IPT ? tile_n - first : (u32)RK_IPT) : 0;
    u64 key[RK_IPT];
    double acc[RK_IPT];
    u32 head_bits = 0, rhead_bits = 0, defer_bits = 0;
    if (mine) {
        u64 prev = s_keys[rk_phys(first)];
#pragma unroll
        for (int j = 0; j < RK_IPT; ++j) {
            key[j] = s_keys[rk_phys(first + j + 1)];
            acc[j] = s_vals[rk_phys(first + j + 1)];
        }
        int cur = -1;  // my open run (index of its head among my entries), -1: none yet
#pragma unroll
        for (int j = 0; j < RK_IPT; ++j) {
            if ((u32)j < mine) {
                const u64 before = j ? key[j - 1] : prev;
                const bool head = (a.policy == POLICY_KEEP_ALL) || (base + first + j == 0) || (key[j] != before);
                if (head) {
                    head_bits |= 1u << j;
                    if (want_rows && ((base + first + j == 0) || (key[j] >> a.bits_lo) != (before >> a.bits_lo)))
                        rhead_bits |= 1u << j;
                    cur = j;
                } else if (cur >= 0) {  // follower of a run that started in my block: fold left to right
                    double v = acc[j];
#pragma unroll
                    for (int h = 0; h < RK_IPT; ++h)
                        if (h == cur) {
                            if (a.policy == POLICY_ADD) acc[h] = __dadd_rn(acc[h], v);
                            else if (a.policy == POLICY_REPLACE) acc[h] = v;
                        }
                }
            }
        }
        // my last run may continue past my block: keep folding (shared memory, then global memory)
        if (cur >= 0 && mine == (u32)RK_IPT && (a.policy == POLICY_ADD || a.policy == POLICY_REPLACE)) {
            const u64 k0 = key[RK_IPT - 1];
            double sum = 0.0;
#pragma unroll
            for (int h = 0; h < RK_IPT; ++h) if (h == cur) sum = acc[h];
            u64 q = base + first + RK_IPT;  // global index of the next entry
            u32 steps = 0;
            while (q < n) {
                const u32 L = (u32)(q - base) + 1;
                u64 kq;
                double vq = 0.0;
                if (L <= (u32)RK_TILE) { kq = s_keys[rk_phys(L)]; vq = s_vals[rk_phys(L)]; }
                else if (L == (u32)RK_TILE + 1) { kq = s_keys[rk_phys(L)]; if (kq == k0) vq = a.vals[q]; }
                else { kq = a.keys[q]; if (kq == k0) vq = a.vals[q]; }
                if (kq != k0) break;
                if (a.policy == POLICY_ADD) sum = __dadd_rn(sum, vq); else sum = vq;
                ++q;
                if (MODE == MODE_CONSOLIDATE && ++steps >= RK_LONG_RUN) { defer_bits |= 1u << cur; break; }
            }
#pragma unroll
            for (int h = 0; h < RK_IPT; ++h) if (h == cur) acc[h] = sum;
        }
    }
    // ---- which heads are emitted ----------------------------------------------------------------------
    u32 emit_bits = head_bits;
    if (MODE == MODE_ESC) {
        const u64 lo_mask = (1ull << a.bits_lo) - 1;
#pragma unroll
        for (int j = 0; j < RK_IPT; ++j)
            if ((head_bits >> j) & 1u) {
                // multiply_sparse.hpp:238: keep iff sum != 0 (NaN kept); masked columns never emitted
                const u32 kcol = (u32)(key[j] & lo_mask);
                if (acc[j] == 0.0 || (a.sk && a.sk[kcol] == 0.0)) emit_bits &= ~(1u << j);
            }
    }
    // ---- slots inside the tile: scan over threads ---------------------------------------------------------
    const u64 my = (u64)__popc(emit_bits) | ((u64)__popc(rhead_bits) << 32);
    const u64 incl = warp_incl_scan(my);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();  // also: every thread is done reading the input tile
    u64 wbefore = 0, total = 0;
#pragma unroll
    for (int w = 0; w < RK_WARPS; ++w) {
        const u64 t = s_warp[w];
        if ((u32)w < warp) wbefore += t;
        total += t;
    }
    const u64 before_me = wbefore + incl - my;
    const u32 tile_out = (u32)(total & 0xffffffffull);
    // ---- warp 0 chains the tiles (decoupled look-back) while the other warps already stage their outputs ----
    if (warp == 0) {
        // look-back value: entries in bits [0,31), rows in bits [31,62)
        const u64 packed = (total & 0xffffffffull) | ((total >> 32) << 31);
        const u64 excl0 = lookback_exclusive(a.state, tile, packed);
        if (lane == 0) {
            s_excl = excl0;
            if (base + RK_TILE >= n) {
                const u64 fin = excl0 + packed;
                const u32 n_out = (u32)(fin & 0x7fffffffull), n_rows = (u32)(fin >> 31);
                *a.out_count = n_out;
                if (want_rows) { *a.row_count = n_rows; a.row_start[n_rows] = n_out; }
            }
        }
    }
    // stage the outputs in shared memory at their tile-local slots (the input tile is dead now)
    {
        u32 slot = (u32)(before_me & 0xffffffffull);
#pragma unroll
        for (int j = 0; j < RK_IPT; ++j) {
            if ((emit_bits >> j) & 1u) {
                double v = acc[j];
                if (MODE == MODE_ESC) {
                    const u64 lo_mask = (1ull << a.bits_lo) - 1;
                    const i32 hi = (i32)(key[j] >> a.bits_lo) + a.row_base;
                    v = __dmul_rn(v, a.C);
                    v = __dmul_rn(v, a.si ? a.si[a.row_ids[hi]] : 1.0);
                    v = __dmul_rn(v, a.sk ? a.sk[(u32)(key[j] & lo_mask)] : 1.0);
                }
                s_keys[slot] = key[j];
                s_vals[slot] = v;
                ++slot;
            }
        }
    }
    __syncthreads();
    const u64 excl_rows = s_excl >> 31;
    const u64 excl = s_excl & 0x7fffffffull;
    // row starts and deferred long runs need the global slot
    if (MODE == MODE_CONSOLIDATE && (rhead_bits | defer_bits)) {
        u32 slot = (u32)(before_me & 0xffffffffull);
        u64 rslot = excl_rows + (before_me >> 32);
#pragma unroll
        for (int j = 0; j < RK_IPT; ++j) {
            if ((emit_bits >> j) & 1u) {
                if ((rhead_bits >> j) & 1u) {
                    a.row_start[rslot] = (u32)(excl + slot);
                    a.row_id[rslot] = (i32)(key[j] >> a.bits_lo);
                    ++rslot;
                }
                if ((defer_bits >> j) & 1u) {
                    u32 t = atomicAdd(a.long_count, 1u);
                    if (t < a.long_cap) { a.long_list[2 * t] = (u32)(excl + slot); a.long_list[2 * t + 1] = (u32)(base + first + j); }
                }
                ++slot;
            }
        }
    }
    // ---- coalesced copy-out, unpacking the key into the two index vectors ----------------------------------
    const u64 lo_mask = (1ull << a.bits_lo) - 1;
    for (u32 t = tid; t < tile_out; t += RK_THREADS) {
        const u64 k = s_keys[t];
        i32 hi = (i32)(k >> a.bits_lo);
        if (MODE == MODE_ESC) hi += a.row_base;
        a.out_hi[excl + t] = hi;
        if (a.out_lo) a.out_lo[excl + t] = (i32)(k & lo_mask);
        a.out_val[excl + t] = s_vals[t];
    }
}

// Runs longer than RK_LONG_RUN: one WARP per run.  The lanes load 32 consecutive values at a time (the next 32 are already in
// flight) and the values are added one after the other, handed over by shuffles -- the reference's left-to-right fold
// (algorithm.hpp:307-310), so these sums are bit-identical to the reference's as well, whatever the run length.
__global__ void __launch_bounds__(256) k_long_runs(ReduceArgs a) {
    const u32 n = *a.n_ptr;
    u32 cnt = *a.long_count;
    if (cnt > a.long_cap) cnt = a.long_cap;
    const u32 lane = lane_id();
    const u32 warps = gridDim.x * (blockDim.x >> 5);
    for (u32 t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < cnt; t += warps) {
        const u32 slot = a.long_list[2 * t], start = a.long_list[2 * t + 1];
        const u64 key = a.keys[start];
        // end of the run: gallop (runs are short next to n), then bisect -- first index with keys[] > key
        u32 lo = start + 1, step = RK_LONG_RUN;
        u32 hi = n;
        for (;;) {
            const u64 probe = (u64)lo + step;
            if (probe >= n) break;
            if (a.keys[probe] > key) { hi = (u32)probe; break; }
            lo = (u32)probe + 1;
            step *= 2;
        }
        while (lo < hi) {
            const u32 mid = lo + (hi - lo) / 2;
            if (a.keys[mid] <= key) lo = mid + 1; else hi = mid;
        }
        const u32 end = lo;
        double r;
        if (a.policy == POLICY_ADD) {
            double sum = 0.0;
            bool first = true;
            u32 i = start;
            double v = (i + lane < end) ? a.vals[i + lane] : 0.0;
            while (i < end) {
                const u32 m = end - i < 32u ? end - i : 32u;
                const u32 inext = i + 32;
                const double vn = (inext < end && inext + lane < end) ? a.vals[inext + lane] : 0.0;
                for (u32 l = 0; l < m; ++l) {  // warp-uniform trip count
                    const double x = __shfl_sync(SPB_FULL_MASK, v, (int)l);
                    if (first) { sum = x; first = false; }   // acc = first value (algorithm.hpp:279)
                    else sum = __dadd_rn(sum, x);
                }
                v = vn;
                i = inext;
            }
            r = sum;
        } else {
            r = (a.policy == POLICY_REPLACE) ? a.vals[end - 1] : a.vals[start];
        }
        if (lane == 0) a.out_val[slot] = r;
    }
}
