// reduce_segsort.cuh -- the in-row column sort (radix_sort.cuh: k_segment_sort_walk) and the duplicate-reduce
// (reduce_by_key.cuh: k_reduce_by_key<MODE_CONSOLIDATE>) in ONE pass over the array.
//
// After the radix passes over the row digits the array is grouped by row, every row still in insertion order.
// k_segment_sort_walk then reads and rewrites the whole array to order each row by column, and the reduce pass
// reads it once more.  Here a reduce tile loads its 2048 entries plus SEG_MAX entries either side, works out where
// every entry of that window goes once its row is sorted (the same count-my-predecessors walk), places the entries
// whose destination lies in the tile (plus SEG_MAX slots behind it, where a duplicate run that starts in the tile
// can still continue) and carries on exactly as k_reduce_by_key does: heads fold their followers left to right
// (the reference's association order, algorithm.hpp:277-313 -- equal keys are adjacent and in insertion order,
// because the walk is stable), outputs are compacted through a decoupled look-back, row starts are emitted.
// One read of the array (16 B per entry) and one write of the result instead of two reads and a write more.
//
// A row of more than SEG_MAX entries cannot be ranked inside a window.  The kernel counts the entries of such rows
// (`seg_long`); when there are any the host discards this launch's output and runs the separate kernels, which have
// a path for long rows (spb_api.cu: sort_reduce).
#pragma once
#include "radix_sort.cuh"
#include "reduce_by_key.cuh"

constexpr int RF_H = SEG_MAX;                                   // halo either side of the tile
constexpr int RF_W = RK_TILE + 2 * RF_H;                        // raw window: entries base - RF_H .. base + RK_TILE + RF_H - 1
constexpr int RF_ITS = (RF_W + RK_THREADS - 1) / RK_THREADS;    // window entries per thread
constexpr int RF_SORTED = RK_TILE + RF_H + 1;                   // sorted slots L = 0 .. RK_TILE + RF_H: entry base - 1 + L
constexpr int RF_SLOTS = RF_SORTED + (RF_SORTED >> 3) + 1;      // with rk_phys's pad word every 8 slots

struct RfSmem {
    u64 raw[RF_W];          // keys of the window as they are in memory (grouped by row, unsorted inside)
    u64 keys[RF_SLOTS];     // keys in sorted order; later: staged output keys
    double vals[RF_SLOTS];  // values in sorted order; later: staged output values
    u64 warp[RK_WARPS];     // per-warp totals: entries | rows << 32
    u64 excl;
    u32 tile;
};

__global__ void __launch_bounds__(RK_THREADS, 4) k_reduce_segsort(ReduceArgs a, u32 *seg_long) {
    extern __shared__ __align__(16) unsigned char rf_smem[];
    RfSmem &s = *reinterpret_cast<RfSmem *>(rf_smem);
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool want_rows = a.row_start != nullptr;
    if (tid == 0) s.tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const u32 tile = s.tile;
    const u32 n = *a.n_ptr;
    const u64 base = (u64)tile * RK_TILE;
    if (base >= n) return;
    const u32 tile_n = (n - base < (u64)RK_TILE) ? (u32)(n - base) : (u32)RK_TILE;

    // ---- the raw window: keys to shared memory, values to registers ------------------------------------------
    double v[RF_ITS];
#pragma unroll
    for (int k = 0; k < RF_ITS; ++k) {
        const u32 q = (u32)k * RK_THREADS + tid;
        const i64 g = (i64)base + (i64)q - RF_H;
        const bool in = q < (u32)RF_W && g >= 0 && g < (i64)n;
        if (q < (u32)RF_W) s.raw[q] = in ? ld_stream_u64(a.keys + g) : ~0ull;  // ~0: belongs to no row
        v[k] = in ? ld_stream_f64(a.vals + g) : 0.0;
    }
    __syncthreads();

    // ---- where every window entry goes once its row is ordered by column ----------------------------------------
    // An entry's place inside its row = the entries of the row that must precede it: smaller column, or equal
    // column and earlier (stable).  A row that reaches the window's edge may continue outside ("cut"): its
    // entries keep their places -- they only ever serve as "a key of another row" to the entries of the tile.
    const u64 lo_mask = (1ull << a.bits_lo) - 1;
    u32 n_long = 0;
#pragma unroll
    for (int k = 0; k < RF_ITS; ++k) {
        const u32 q = (u32)k * RK_THREADS + tid;
        if (q >= (u32)RF_W) continue;
        const u64 key = s.raw[q];
        if (key == ~0ull) continue;
        const u64 row = key >> a.bits_lo, col = key & lo_mask;
        u32 b = 0, f = 0, before = 0;
        bool cut = false;
        for (; b < (u32)SEG_MAX; ++b) {   // earlier entries of my row
            if (q < 1 + b) { cut = true; break; }
            const u64 kk = s.raw[q - 1 - b];
            if ((kk >> a.bits_lo) != row) break;
            before += (kk & lo_mask) <= col;
        }
        for (; f < (u32)SEG_MAX; ++f) {   // later entries of my row
            if (q + 1 + f >= (u32)RF_W) { cut = true; break; }
            const u64 kk = s.raw[q + 1 + f];
            if ((kk >> a.bits_lo) != row) break;
            before += (kk & lo_mask) < col;
        }
        const bool is_long = b == (u32)SEG_MAX || f == (u32)SEG_MAX || b + f + 1 > (u32)SEG_MAX;
        // entries of the tile itself are SEG_MAX away from both edges: for them "long" means long
        if (q >= (u32)RF_H && q < (u32)(RF_H + RK_TILE)) n_long += is_long;
        const u32 place = (cut || is_long) ? q : q - b + before;   // window position after the row sort
        if (place + 1 >= (u32)RF_H && place + 1 < (u32)(RF_H + RF_SORTED)) {
            const u32 L = place + 1 - (u32)RF_H;
            s.keys[rk_phys(L)] = key;
            s.vals[rk_phys(L)] = v[k];
        }
    }
    n_long = __reduce_add_sync(SPB_FULL_MASK, n_long);
    if (n_long && lane == 0) atomicAdd(seg_long, n_long);
    __syncthreads();

    // ---- from here on: k_reduce_by_key<MODE_CONSOLIDATE> over the sorted tile in shared memory ----------------------
    const u32 first = tid * RK_IPT;  // tile-local index of my first entry
    const u32 mine = first < tile_n ? (tile_n - first < (u32)RK_IPT ? tile_n - first : (u32)RK_IPT) : 0;
    u64 key[RK_IPT];
    double acc[RK_IPT];
    u32 head_bits = 0, rhead_bits = 0;
    if (mine) {
        const u64 prev = s.keys[rk_phys(first)];
#pragma unroll
        for (int j = 0; j < RK_IPT; ++j) {
            key[j] = s.keys[rk_phys(first + j + 1)];
            acc[j] = s.vals[rk_phys(first + j + 1)];
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
                    const double vj = acc[j];
#pragma unroll
                    for (int h = 0; h < RK_IPT; ++h)
                        if (h == cur) {
                            if (a.policy == POLICY_ADD) acc[h] = __dadd_rn(acc[h], vj);
                            else if (a.policy == POLICY_REPLACE) acc[h] = vj;
                        }
                }
            }
        }
        // my last run may continue past my block -- and past the tile: equal keys share a row, so the run ends
        // within SEG_MAX entries, all of which were placed above
        if (cur >= 0 && mine == (u32)RK_IPT && (a.policy == POLICY_ADD || a.policy == POLICY_REPLACE)) {
            const u64 k0 = key[RK_IPT - 1];
            double sum = 0.0;
#pragma unroll
            for (int h = 0; h < RK_IPT; ++h) if (h == cur) sum = acc[h];
            u64 q = base + first + RK_IPT;  // global index of the next entry
            while (q < n) {
                const u32 L = (u32)(q - base) + 1;
                if (L >= (u32)RF_SORTED) break;  // only inside a row longer than SEG_MAX (this launch is discarded then)
                if (s.keys[rk_phys(L)] != k0) break;
                const double vq = s.vals[rk_phys(L)];
                if (a.policy == POLICY_ADD) sum = __dadd_rn(sum, vq); else sum = vq;
                ++q;
            }
#pragma unroll
            for (int h = 0; h < RK_IPT; ++h) if (h == cur) acc[h] = sum;
        }
    }
    // ---- slots inside the tile: scan over threads ---------------------------------------------------------
    const u32 emit_bits = head_bits;
    const u64 my = (u64)__popc(emit_bits) | ((u64)__popc(rhead_bits) << 32);
    const u64 incl = warp_incl_scan(my);
    if (lane == 31) s.warp[warp] = incl;
    __syncthreads();  // also: every thread is done reading the sorted tile
    u64 wbefore = 0, total = 0;
#pragma unroll
    for (int w = 0; w < RK_WARPS; ++w) {
        const u64 t = s.warp[w];
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
            s.excl = excl0;
            if (base + RK_TILE >= n) {
                const u64 fin = excl0 + packed;
                const u32 n_out = (u32)(fin & 0x7fffffffull), n_rows = (u32)(fin >> 31);
                *a.out_count = n_out;
                if (want_rows) { *a.row_count = n_rows; a.row_start[n_rows] = n_out; }
            }
        }
    }
    // stage the outputs at their tile-local slots (the sorted tile is dead now)
    {
        u32 slot = (u32)(before_me & 0xffffffffull);
#pragma unroll
        for (int j = 0; j < RK_IPT; ++j) {
            if ((emit_bits >> j) & 1u) {
                s.keys[slot] = key[j];
                s.vals[slot] = acc[j];
                ++slot;
            }
        }
    }
    __syncthreads();
    const u64 excl_rows = s.excl >> 31;
    const u64 excl = s.excl & 0x7fffffffull;
    if (rhead_bits) {  // row starts need the global slot
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
                ++slot;
            }
        }
    }
    // ---- coalesced copy-out, unpacking the key into the two index vectors ----------------------------------
    for (u32 t = tid; t < tile_out; t += RK_THREADS) {
        const u64 k = s.keys[t];
        a.out_hi[excl + t] = (i32)(k >> a.bits_lo);
        if (a.out_lo) a.out_lo[excl + t] = (i32)(k & lo_mask);
        a.out_val[excl + t] = s.vals[t];
    }
}
