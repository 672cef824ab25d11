// reduce_by_key.cuh -- segmented duplicate-reduce + compaction over sorted packed keys.
//
// Replaces the merge loop of spsparse::consolidate (reference slib/spsparse/algorithm.hpp:277-313):
// equal keys are adjacent after the stable sort and still in insertion order, so the head of
// every run folds its followers left to right -- acc = v0; acc += v1; ... -- which is the
// reference's association order, hence bit-identical sums.  Runs longer than RK_LONG_RUN are
// finished by k_long_runs with a block-wide tree (same value to ~1 ulp * log n, documented).
// Output slots come from a single-pass decoupled look-back over the tiles' head counts.
//
// The same kernel is the "compress" step of expand-sort-compress in multiply (MODE_ESC): there
// the folded sum is the dot product of multiply_sparse.hpp:219-236, dropped when exactly zero
// (:238), scaled as ((sum*C)*a_scale)*b_scale (:242).
#pragma once
#include "scan.cuh"

constexpr int RK_THREADS = 256;
constexpr int RK_IPT = 8;
constexpr int RK_TILE = RK_THREADS * RK_IPT;
constexpr int RK_WARPS = RK_THREADS / 32;
constexpr u32 RK_LONG_RUN = 4096;

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

template <int MODE>
__global__ void __launch_bounds__(RK_THREADS) k_reduce_by_key(ReduceArgs a) {
    __shared__ u64 s_keys[RK_TILE + 1];
    __shared__ double s_vals[RK_TILE];
    __shared__ u64 s_part[RK_IPT * RK_WARPS];  // low 32 bits: entries emitted, high 32 bits: row heads
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

    // tile -> shared memory; s_keys[0] is the predecessor of the tile's first entry
#pragma unroll
    for (int k = 0; k < RK_IPT; ++k) {
        u32 p = (u32)k * RK_THREADS + tid;
        u64 i = base + p;
        if (i < n) {
            s_keys[p + 1] = ld_stream_u64(a.keys + i);
            s_vals[p] = ld_stream_f64(a.vals + i);
        }
    }
    if (tid == 0) s_keys[0] = base ? a.keys[base - 1] : 0;
    __syncthreads();

    const u32 lt = lanemask_lt();
    double acc[RK_IPT];
    u32 emit_bits = 0, defer_bits = 0, rhead_bits = 0, rank_in_warp[RK_IPT];
    unsigned char rrank_in_warp[RK_IPT];
#pragma unroll
    for (int k = 0; k < RK_IPT; ++k) {
        u32 p = (u32)k * RK_THREADS + tid;
        u64 i = base + p;
        bool head = false;
        double sum = 0.0;
        if (i < n) {
            const u64 key = s_keys[p + 1];
            head = (a.policy == POLICY_KEEP_ALL) || (i == 0) || (key != s_keys[p]);
            if (head) {
                sum = s_vals[p];
                if (a.policy == POLICY_ADD || a.policy == POLICY_REPLACE) {
                    // fold the followers of this run, left to right
                    u64 q = i + 1;
                    u32 steps = 0;
                    bool more = true;
                    while (more && q < n) {
                        u32 pq = (u32)(q - base);
                        u64 kq;
                        double vq;
                        if (pq < RK_TILE) { kq = s_keys[pq + 1]; vq = s_vals[pq]; }
                        else { kq = a.keys[q]; vq = (kq == key) ? a.vals[q] : 0.0; }
                        if (kq != key) break;
                        if (a.policy == POLICY_ADD) sum = __dadd_rn(sum, vq);
                        else if (a.policy == POLICY_REPLACE) sum = vq;
                        ++q;
                        if (MODE == MODE_CONSOLIDATE && ++steps >= RK_LONG_RUN) {
                            defer_bits |= 1u << k;  // finished by k_long_runs
                            more = false;
                        }
                    }
                }
            }
        }
        bool emit = head;
        if (MODE == MODE_ESC && head) {
            // multiply_sparse.hpp:238: keep iff sum != 0 (NaN kept); masked columns never emitted
            u32 kcol = (u32)(s_keys[p + 1] & ((1ull << a.bits_lo) - 1));
            if (sum == 0.0 || (a.sk && a.sk[kcol] == 0.0)) emit = false;
        }
        acc[k] = sum;
        u32 b = __ballot_sync(SPB_FULL_MASK, emit);
        rank_in_warp[k] = __popc(b & lt);
        if (emit) emit_bits |= 1u << k;
        u32 rb = 0;
        if (want_rows) {
            // first output entry of a new leading index: its predecessor (always a different key) has another hi
            bool rhead = emit && (i == 0 || (s_keys[p + 1] >> a.bits_lo) != (s_keys[p] >> a.bits_lo));
            rb = __ballot_sync(SPB_FULL_MASK, rhead);
            rrank_in_warp[k] = (unsigned char)__popc(rb & lt);
            if (rhead) rhead_bits |= 1u << k;
        }
        if (lane == 0) s_part[k * RK_WARPS + warp] = (u64)__popc(b) | ((u64)__popc(rb) << 32);
    }
    __syncthreads();
    if (warp == 0) {
        u64 x = s_part[2 * lane], y = s_part[2 * lane + 1];
        u64 s = warp_incl_scan(x + y);  // both halves at once: neither can carry into the other
        u64 total = __shfl_sync(SPB_FULL_MASK, s, 31);
        s_part[2 * lane] = s - x - y;
        s_part[2 * lane + 1] = s - y;
        // look-back value: entries in bits [0,31), rows in bits [31,62)
        u64 packed = (total & 0xffffffffull) | ((total >> 32) << 31);
        u64 excl = lookback_exclusive(a.state, tile, packed);
        if (lane == 0) {
            s_excl = excl;
            if (base + RK_TILE >= n) {
                u64 fin = excl + packed;
                u32 n_out = (u32)(fin & 0x7fffffffull), n_rows = (u32)(fin >> 31);
                *a.out_count = n_out;
                if (want_rows) { *a.row_count = n_rows; a.row_start[n_rows] = n_out; }
            }
        }
    }
    __syncthreads();
    const u64 excl_rows = s_excl >> 31;
    const u64 excl = s_excl & 0x7fffffffull;
    const u64 lo_mask = (1ull << a.bits_lo) - 1;
#pragma unroll
    for (int k = 0; k < RK_IPT; ++k) {
        if (!((emit_bits >> k) & 1u)) continue;
        u32 p = (u32)k * RK_THREADS + tid;
        u64 key = s_keys[p + 1];
        u64 slot = excl + (u32)s_part[k * RK_WARPS + warp] + rank_in_warp[k];
        i32 hi = (i32)(key >> a.bits_lo), lo = (i32)(key & lo_mask);
        if (want_rows && ((rhead_bits >> k) & 1u)) {
            u64 rslot = excl_rows + (u32)(s_part[k * RK_WARPS + warp] >> 32) + rrank_in_warp[k];
            a.row_start[rslot] = (u32)slot;
            a.row_id[rslot] = hi;
        }
        if (MODE == MODE_ESC) hi += a.row_base;
        if (MODE == MODE_CONSOLIDATE) {
            a.out_hi[slot] = hi;
            if (a.out_lo) a.out_lo[slot] = lo;
            a.out_val[slot] = acc[k];
            if ((defer_bits >> k) & 1u) {
                u32 t = atomicAdd(a.long_count, 1u);
                if (t < a.long_cap) { a.long_list[2 * t] = (u32)slot; a.long_list[2 * t + 1] = (u32)(base + p); }
            }
        } else {
            double v = __dmul_rn(acc[k], a.C);
            v = __dmul_rn(v, a.si ? a.si[a.row_ids[hi]] : 1.0);
            v = __dmul_rn(v, a.sk ? a.sk[lo] : 1.0);
            a.out_hi[slot] = hi;  // compressed row number; mapped to i when copied into C
            a.out_lo[slot] = lo;
            a.out_val[slot] = v;
        }
    }
}

// Runs longer than RK_LONG_RUN: one block per run, binary search for its end, tree reduction.
__global__ void __launch_bounds__(256) k_long_runs(ReduceArgs a) {
    __shared__ double s_red[256];
    const u32 n = *a.n_ptr;
    u32 cnt = *a.long_count;
    if (cnt > a.long_cap) cnt = a.long_cap;
    for (u32 t = blockIdx.x; t < cnt; t += gridDim.x) {
        const u32 slot = a.long_list[2 * t], start = a.long_list[2 * t + 1];
        const u64 key = a.keys[start];
        u32 lo = start, hi = n;  // first index with keys[] > key
        while (lo < hi) {
            u32 mid = lo + (hi - lo) / 2;
            if (a.keys[mid] <= key) lo = mid + 1; else hi = mid;
        }
        const u32 end = lo;
        double r;
        if (a.policy == POLICY_ADD) {
            double s = 0.0;
            for (u32 i = start + threadIdx.x; i < end; i += blockDim.x) s += a.vals[i];
            s_red[threadIdx.x] = s;
            __syncthreads();
            for (u32 o = 128; o > 0; o >>= 1) {
                if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
                __syncthreads();
            }
            r = s_red[0];
        } else {
            r = (a.policy == POLICY_REPLACE) ? a.vals[end - 1] : a.vals[start];
        }
        if (threadIdx.x == 0) a.out_val[slot] = r;
        __syncthreads();
    }
}
