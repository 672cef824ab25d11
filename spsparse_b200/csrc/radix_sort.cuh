// radix_sort.cuh -- stable LSD radix sort of packed (hi,lo) index keys with an fp64 payload.
//
// Replaces sorted_permutation + CmpIndex + std::stable_sort (reference slib/spsparse/
// algorithm.hpp:375-427).  One-sweep organisation: a single histogram kernel counts every 8-bit
// digit of every pass up front; each pass then reads its input once and writes it once, ranking
// items inside a 4096-entry tile with a warp-level ballot multi-split and chaining the tiles'
// per-digit counts with a decoupled look-back.  Pass 0 reads the caller's struct-of-arrays COO
// directly, packs (hi << bits_lo) | lo on the fly and applies consolidate()'s input drop rule
// (algorithm.hpp:272-275, 284-292), so filtered entries never enter the sort.
#pragma once
#include "common.cuh"

constexpr int RS_THREADS = 256;
#ifndef RS_IPT_N
#define RS_IPT_N 16      // entries per thread of a radix-pass tile (same-box A/B builds: -DRS_IPT_N=10 -DRS_MIN_BLOCKS=4)
#endif
#ifndef RS_MIN_BLOCKS
#define RS_MIN_BLOCKS 3
#endif
constexpr int RS_IPT = RS_IPT_N;
constexpr int RS_HB = RS_IPT / 2;   // pass 0 without bulk copies: two batches of loads
constexpr int RS_TILE = RS_THREADS * RS_IPT;  // 4096 entries per tile
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_RADIX_BITS = 8;
constexpr int RS_RADIX = 1 << RS_RADIX_BITS;
constexpr int RS_NB = RS_RADIX + 1;  // +1: bucket of filtered / out-of-range items (never written)
constexpr int RS_MAX_PASSES = 8;
#ifndef RS_LOOKBACK
#define RS_LOOKBACK 8  // predecessors read per look-back round trip
#endif
constexpr size_t RS_SMEM_BYTES =
    (size_t)RS_TILE * 16 + (size_t)RS_WARPS * RS_NB * sizeof(u32) + 64;

// Look-back word of the sort: 0 = not published yet; bit 31 set = inclusive prefix of the tile in bits 30:0; otherwise the
// tile's own count + 1 (a tile holds RS_TILE entries, so this never reaches bit 31).  Prefixes up to 2^31 - 1: the
// reference's own limit on the entries of one array (algorithm.hpp:419-421).
#define RS_WORD_AGG(c) ((c) + 1u)
#define RS_WORD_INCL(c) ((c) | 0x80000000u)
#define RS_READY(w) ((w) != 0u)
#define RS_IS_INCL(w) (((w) >> 31) != 0u)
#define RS_VALUE(w) (RS_IS_INCL(w) ? ((w) & 0x7fffffffu) : (w) - 1u)

// Where pass 0 reads from, and which inputs it drops.
struct SortInput {
    const i32 *hi;      // index vector of the leading sort dimension
    const i32 *lo;      // second sort dimension, or nullptr (rank 1)
    const double *val;
    u32 n;
    int bits_lo;        // key = (hi << bits_lo) | lo
    u32 extent_hi, extent_lo;  // bounds check (VectorCooArray::add, VectorCooArray.hpp:245-262)
    int drop_zero;      // drop val == 0 inputs (always, for consolidate)
    int drop_nan;       // zero_nan: drop NaN inputs that precede the first kept entry of the
                        // REFERENCE's sorted sequence (the "leading run", algorithm.hpp:272-275)
    const i32 *ref_hi;  // key in the reference's sort order (may differ from ours when multiply
    const i32 *ref_lo;  // needs B bucketed by its inner index)
    int ref_bits_lo;
    const u64 *first_kept;  // [0] smallest reference key among non-none inputs, [1] its smallest position
};

__device__ __forceinline__ u64 pack_key(i32 hi, i32 lo, int bits_lo) {
    return ((u64)(u32)hi << bits_lo) | (u64)(u32)lo;
}

__device__ __forceinline__ bool input_kept(const SortInput &in, u32 i, double v) {
    if (in.drop_zero && v == 0.0) return false;
    if (in.drop_nan && isnan(v)) {
        u64 rk = pack_key(in.ref_hi[i], in.ref_lo ? in.ref_lo[i] : 0, in.ref_bits_lo);
        u64 kmin = in.first_kept[0];
        if (rk < kmin || (rk == kmin && (u64)i < in.first_kept[1])) return false;
    }
    return true;
}

// ---- zero_nan support: locate the first kept entry of the reference's sorted sequence ---------
__global__ void k_first_kept_key(SortInput in, u64 *first_kept) {
    u64 best = ~0ull;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < in.n; i += (u64)gridDim.x * blockDim.x) {
        double v = in.val[i];
        if (v == 0.0 || isnan(v)) continue;
        u64 rk = pack_key(in.ref_hi[i], in.ref_lo ? in.ref_lo[i] : 0, in.ref_bits_lo);
        best = rk < best ? rk : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u64 t = __shfl_xor_sync(SPB_FULL_MASK, best, o);
        best = t < best ? t : best;
    }
    if (lane_id() == 0 && best != ~0ull) atomicMin((ull *)&first_kept[0], (ull)best);
}
__global__ void k_first_kept_pos(SortInput in, u64 *first_kept) {
    const u64 kmin = first_kept[0];
    u64 best = ~0ull;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < in.n; i += (u64)gridDim.x * blockDim.x) {
        double v = in.val[i];
        if (v == 0.0 || isnan(v)) continue;
        u64 rk = pack_key(in.ref_hi[i], in.ref_lo ? in.ref_lo[i] : 0, in.ref_bits_lo);
        if (rk == kmin) best = i < best ? i : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u64 t = __shfl_xor_sync(SPB_FULL_MASK, best, o);
        best = t < best ? t : best;
    }
    if (lane_id() == 0 && best != ~0ull) atomicMin((ull *)&first_kept[1], (ull)best);
}

// ---- histogram of every digit of every pass (one read of the index vectors + values) -----------
// counters[0] = kept entries, counters[1] = 1 if an index was out of bounds.
__global__ void __launch_bounds__(512) k_sort_hist(SortInput in, int passes, int shift0, u32 *hist, u32 *counters) {
    __shared__ u32 s_h[RS_MAX_PASSES * RS_RADIX];
    for (int t = threadIdx.x; t < passes * RS_RADIX; t += blockDim.x) s_h[t] = 0;
    __syncthreads();
    u32 kept = 0, oob = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    constexpr int HU = 4;  // entries per thread per trip: all of their loads are issued before the first counter is touched
    for (u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; i0 < in.n; i0 += HU * stride) {
        i32 hi[HU], lo[HU];
        double v[HU];
#pragma unroll
        for (int u = 0; u < HU; ++u) {
            const u64 i = i0 + u * stride;
            const u64 ic = i < in.n ? i : i0;
            hi[u] = ld_stream_i32(in.hi + ic);
            lo[u] = in.lo ? ld_stream_i32(in.lo + ic) : 0;
            v[u] = ld_stream_f64(in.val + ic);
        }
#pragma unroll
        for (int u = 0; u < HU; ++u) {
            const u64 i = i0 + u * stride;
            if (i >= in.n) break;
            if ((u32)hi[u] >= in.extent_hi || (u32)lo[u] >= in.extent_lo) { oob = 1; continue; }
            if (!input_kept(in, (u32)i, v[u])) continue;
            ++kept;
            u64 key = pack_key(hi[u], lo[u], in.bits_lo) >> shift0;  // passes cover the key bits from shift0 upwards
            for (int p = 0; p < passes; ++p) {
                atomicAdd(&s_h[p * RS_RADIX + (u32)(key & (RS_RADIX - 1))], 1u);
                key >>= RS_RADIX_BITS;
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < passes * RS_RADIX; t += blockDim.x)
        if (s_h[t]) atomicAdd(&hist[t], s_h[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(SPB_FULL_MASK, kept, o);
    if (lane_id() == 0 && kept) atomicAdd(&counters[0], kept);
    if (oob) counters[1] = 1;
}

// histogram of already packed keys (expand-sort-compress path)
__global__ void __launch_bounds__(512) k_keys_hist(const u64 *__restrict__ keys, u64 n, int passes, u32 *hist) {
    __shared__ u32 s_h[RS_MAX_PASSES * RS_RADIX];
    for (int t = threadIdx.x; t < passes * RS_RADIX; t += blockDim.x) s_h[t] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u64 key = ld_stream_u64(keys + i);
        for (int p = 0; p < passes; ++p) {
            atomicAdd(&s_h[p * RS_RADIX + (u32)(key & (RS_RADIX - 1))], 1u);
            key >>= RS_RADIX_BITS;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < passes * RS_RADIX; t += blockDim.x)
        if (s_h[t]) atomicAdd(&hist[t], s_h[t]);
}

// exclusive scan of each pass's 256 counts -> first output slot of each digit (in place)
__global__ void __launch_bounds__(RS_RADIX) k_bucket_starts(u32 *hist) {
    __shared__ u32 s_w[RS_RADIX / 32];
    u32 *h = hist + blockIdx.x * RS_RADIX;
    u32 c = h[threadIdx.x];
    u32 incl = warp_incl_scan(c);
    if (lane_id() == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 add = 0;
    for (u32 w = 0; w < (threadIdx.x >> 5); ++w) add += s_w[w];
    h[threadIdx.x] = add + incl - c;
}

struct PassArgs {
    const u64 *keys_in;      // passes >= 1
    const double *vals_in;
    u64 *keys_out;
    double *vals_out;
    const u32 *n_ptr;        // device-resident entry count for passes >= 1 (n_kept)
    const u32 *bucket_start; // [256], this pass
    u32 *lookback;           // [tiles][256], zeroed
    u32 *ticket;             // zeroed
    int shift;               // digit = (key >> shift) & 255
    int rank_mode;           // 0 ballots (default), 1 match.any (kept for A/B measurements; SPB_RANK_MODE)
};

// BULK (passes after the first): the tile's keys and values are brought into shared memory by two bulk asynchronous copies
// (cp.async.bulk + mbarrier; one thread issues them, 64 KB in flight per block at no register cost) instead of 16 + 16 loads
// per thread; the values arrive while the keys are being ranked.
template <bool PASS0, bool BULK = false>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_BLOCKS) k_radix_pass(PassArgs a, SortInput in) {
    __shared__ __align__(8) u64 s_mbar[2];
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u64 *s_keys = reinterpret_cast<u64 *>(smem_raw);
    double *s_vals = reinterpret_cast<double *>(s_keys + RS_TILE);
    u32 *s_wcnt = reinterpret_cast<u32 *>(s_vals + RS_TILE);  // [RS_WARPS][RS_NB]; later: global bases
    u32 *s_misc = s_wcnt + RS_WARPS * RS_NB;                 // [0] tile, [1..8] warp sums

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_misc[0] = atomicAdd(a.ticket, 1u);
        if (BULK) { mbar_init(&s_mbar[0], 1); mbar_init(&s_mbar[1], 1); mbar_fence_init(); }
    }
    for (int t = tid; t < RS_WARPS * RS_NB; t += RS_THREADS) s_wcnt[t] = 0;
    __syncthreads();
    const u32 tile = s_misc[0];
    const u32 n = PASS0 ? in.n : *a.n_ptr;
    const u64 tile_base = (u64)tile * RS_TILE;
    if (tile_base >= n) return;
    if (BULK && PASS0 && tid == 0) {
        // pass 0, FULL tiles only (the host runs the last, partial tile through the non-bulk kernel: the caller's arrays end
        // where they end): row indices into the first half of the key buffer, column indices into the second, values in place
        i32 *s_idx = reinterpret_cast<i32 *>(s_keys);
        mbar_expect_tx(&s_mbar[0], (in.lo ? 2u : 1u) * RS_TILE * 4u);
        bulk_load(s_idx, in.hi + tile_base, RS_TILE * 4u, &s_mbar[0]);
        if (in.lo) bulk_load(s_idx + RS_TILE, in.lo + tile_base, RS_TILE * 4u, &s_mbar[0]);
        mbar_expect_tx(&s_mbar[1], RS_TILE * 8u);
        bulk_load(s_vals, in.val + tile_base, RS_TILE * 8u, &s_mbar[1]);
    }
    if (BULK && !PASS0 && tid == 0) {
        // (a last tile with an odd number of entries reads 8 bytes past entry n - 1: inside the buffer, whose size is rounded up)
        const u32 valid = n - tile_base < (u64)RS_TILE ? (u32)(n - tile_base) : (u32)RS_TILE;
        const u32 bytes = (valid * 8u + 15u) & ~15u;
        mbar_expect_tx(&s_mbar[0], bytes);
        bulk_load(s_keys, a.keys_in + tile_base, bytes, &s_mbar[0]);
        mbar_expect_tx(&s_mbar[1], bytes);
        bulk_load(s_vals, a.vals_in + tile_base, bytes, &s_mbar[1]);
    }

    // ---- load (warp-striped: item order inside the tile is (warp, k, lane)) ------------------
    const u64 wbase = tile_base + (u64)warp * (32 * RS_IPT) + lane;
    u64 key[RS_IPT];
    u32 valid_bits = 0;
    if (PASS0 && BULK) {
        mbar_wait(&s_mbar[0], 0);
        mbar_wait(&s_mbar[1], 0);
        const i32 *s_idx = reinterpret_cast<const i32 *>(s_keys);
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k) {
            const u32 q = warp * (32 * RS_IPT) + k * 32 + lane;
            const i32 hi = s_idx[q], lo = in.lo ? s_idx[RS_TILE + q] : 0;
            const bool ok = ((u32)hi < in.extent_hi) && ((u32)lo < in.extent_lo) && input_kept(in, (u32)(tile_base + q), s_vals[q]);
            key[k] = pack_key(hi, lo, in.bits_lo);
            valid_bits |= (ok ? 1u : 0u) << k;
        }
    } else if (PASS0) {
        // two batches of 8 so that at most 8 x (hi, lo, val) loads are in flight per thread; the value
        // is only tested here (drop rule) and re-read from L2 when it is staged below
#pragma unroll
        for (int h = 0; h < RS_IPT; h += RS_HB) {
            i32 hi[RS_HB], lo[RS_HB];
            double v[RS_HB];
#pragma unroll
            for (int k = 0; k < RS_HB; ++k) {
                u64 i = wbase + (u64)(h + k) * 32;
                u64 ic = i < n ? i : (u64)n - 1;  // clamp: loads stay unconditional
                hi[k] = ld_stream_i32(in.hi + ic);
                lo[k] = in.lo ? ld_stream_i32(in.lo + ic) : 0;
                v[k] = in.val[ic];
            }
#pragma unroll
            for (int k = 0; k < RS_HB; ++k) {
                u64 i = wbase + (u64)(h + k) * 32;
                bool ok = (i < n) && ((u32)hi[k] < in.extent_hi) && ((u32)lo[k] < in.extent_lo) &&
                          input_kept(in, (u32)i, v[k]);
                key[h + k] = pack_key(hi[k], lo[k], in.bits_lo);
                valid_bits |= (ok ? 1u : 0u) << (h + k);
            }
        }
    } else if (BULK) {
        mbar_wait(&s_mbar[0], 0);
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k) {
            const bool ok = wbase + (u64)k * 32 < n;
            key[k] = ok ? s_keys[warp * (32 * RS_IPT) + k * 32 + lane] : 0;
            valid_bits |= (ok ? 1u : 0u) << k;
        }
    } else {
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k) {
            u64 i = wbase + (u64)k * 32;
            bool ok = i < n;
            key[k] = ok ? ld_stream_u64(a.keys_in + i) : 0;
            valid_bits |= (ok ? 1u : 0u) << k;
        }
        // the values are staged after the ranking: pull their lines into L2 now (no registers held)
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k) {
            u64 i = wbase + (u64)k * 32;
            if (i < n && (lane & 15) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.vals_in + i));
        }
    }

    // ---- rank inside the warp: ballot multi-split, warp-private digit counters ----------------
    // (eight VOTEs + logic per item; MATCH.ANY costs ~64 ADU cycles per warp on sm_100 and capped the
    // pass at ~45% of HBM bandwidth -- see profiles/r01_radix_pass_notes.md)
    u32 *mycnt = s_wcnt + warp * RS_NB;
    const u32 lt = lanemask_lt();
    unsigned short pos[RS_IPT];
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        const bool ok = (valid_bits >> k) & 1u;
        const u32 d = (u32)((key[k] >> a.shift) & (RS_RADIX - 1));
        u32 peers;
        if (a.rank_mode == 0) {
            peers = __ballot_sync(SPB_FULL_MASK, ok);
#pragma unroll
            for (int b = 0; b < RS_RADIX_BITS; ++b) {
                // s = all-ones if bit b of the digit is set, else 0 (one sign-extending bit-field extract);
                // lanes that agree with me on this bit are ~(m ^ s): a single three-input logic op
                int sgn;
                asm("bfe.s32 %0, %1, %2, 1;" : "=r"(sgn) : "r"(d), "r"(b));
                const u32 m = __ballot_sync(SPB_FULL_MASK, sgn != 0);
                peers &= ~(m ^ (u32)sgn);
            }
        } else {
            peers = __match_any_sync(SPB_FULL_MASK, ok ? d : (u32)RS_RADIX);
        }
        u32 leader = ok ? (u32)(__ffs(peers) - 1) : lane;
        u32 before = 0;
        if (ok && lane == leader) {
            before = mycnt[d];
            mycnt[d] = before + __popc(peers);
        }
        before = __shfl_sync(SPB_FULL_MASK, before, leader);
        pos[k] = (unsigned short)(before + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit (thread d): offsets of each warp, tile total, start inside the tile --------
    u32 total = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) total += s_wcnt[w * RS_NB + tid];
    // publish this tile's count of digit `tid` as early as possible
    u32 *lb = a.lookback + (u64)tile * RS_RADIX + tid;
    st_relaxed_u32(lb, tile == 0 ? RS_WORD_INCL(total) : RS_WORD_AGG(total));
    u32 incl = warp_incl_scan(total);
    if (lane == 31) s_misc[1 + warp] = incl;
    __syncthreads();
    u32 lstart = incl - total;
    for (u32 w = 0; w < warp; ++w) lstart += s_misc[1 + w];
    const u32 nvalid_if_last = lstart + total;  // meaningful for tid == 255
    {
        u32 run = lstart;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            u32 c = s_wcnt[w * RS_NB + tid];
            s_wcnt[w * RS_NB + tid] = run;
            run += c;
        }
    }
    if (tid == RS_RADIX - 1) s_misc[12] = nvalid_if_last;
    __syncthreads();
    const u32 nvalid = s_misc[12];

    // ---- stage keys and values in digit order in shared memory --------------------------------
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        if ((valid_bits >> k) & 1u) {
            u32 d = (u32)((key[k] >> a.shift) & (RS_RADIX - 1));
            u32 p = mycnt[d] + pos[k];
            pos[k] = (unsigned short)p;
            s_keys[p] = key[k];
        }
    }
    {
        const double *vsrc = PASS0 ? in.val : a.vals_in;
        double v[RS_IPT];
        if (BULK) {
            // the values sit in s_vals in input order: into registers, then (everybody has read) back in digit order
            mbar_wait(&s_mbar[1], 0);
#pragma unroll
            for (int k = 0; k < RS_IPT; ++k) v[k] = s_vals[warp * (32 * RS_IPT) + k * 32 + lane];
            __syncthreads();
        } else {
#pragma unroll
            for (int k = 0; k < RS_IPT; ++k) {
                u64 i = wbase + (u64)k * 32;
                v[k] = ((valid_bits >> k) & 1u) ? ld_stream_f64(vsrc + i) : 0.0;
            }
        }
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k)
            if ((valid_bits >> k) & 1u) s_vals[pos[k]] = v[k];
    }

    // ---- decoupled look-back, one digit per thread ----------------------------------------------
    // Four predecessors are read at once: the walk back to the nearest tile with an inclusive prefix
    // costs one L2 round trip per four hops instead of one per hop.
    u32 excl = 0;
    if (tile > 0) {
        i64 p = (i64)tile - 1;
        bool done = false;
        while (!done) {
            u32 w[RS_LOOKBACK];
#pragma unroll
            for (int u = 0; u < RS_LOOKBACK; ++u)
                w[u] = (p - u >= 0) ? ld_relaxed_u32(a.lookback + (u64)(p - u) * RS_RADIX + tid) : RS_WORD_INCL(0u);
#pragma unroll
            for (int u = 0; u < RS_LOOKBACK; ++u) {
                if (done) break;
                while (!RS_READY(w[u])) w[u] = ld_relaxed_u32(a.lookback + (u64)(p - u) * RS_RADIX + tid);
                excl += RS_VALUE(w[u]);
                if (RS_IS_INCL(w[u])) done = true;
            }
            p -= RS_LOOKBACK;
        }
        st_relaxed_u32(lb, RS_WORD_INCL(excl + total));
    }
    const u32 gbase = a.bucket_start[tid] + excl - lstart;  // global slot = gbase + position in tile
    __syncthreads();  // all staging done, s_wcnt free
    s_wcnt[tid] = gbase;
    __syncthreads();

    // ---- write out: consecutive threads -> consecutive staged items -> runs of consecutive slots -
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        u32 p = (u32)k * RS_THREADS + tid;
        if (p < nvalid) {
            u64 kk = s_keys[p];
            u32 dst = s_wcnt[(u32)((kk >> a.shift) & (RS_RADIX - 1))] + p;
            a.keys_out[dst] = kk;
            a.vals_out[dst] = s_vals[p];
        }
    }
}

// ---- column order inside the rows -----------------------------------------------------------------------------
// When the column part of the key is two or more digits wide, the radix passes sort by the ROW part only (stable:
// the entries of a row stay in insertion order) and this kernel then orders every row's entries by column: entry e
// looks at the other entries of its row, counts those that must precede it (smaller column, or equal column and
// earlier = stable) and moves to row start + that count.  Rows are short in the matrices this library is for (a
// handful of entries; the walk is bounded by SEG_MAX either way), so this is one read and one write of the array in
// place of (column bits / 8) full passes.  Entries of rows longer than SEG_MAX keep their place and are counted
// (and flagged when `flags` is given): the host sorts exactly those entries by their full key afterwards.
constexpr int SG_THREADS = 256;
constexpr int SG_IPT = 8;
constexpr int SG_TILE = SG_THREADS * SG_IPT;
constexpr int SEG_MAX = 64;

// By-product for the reduce pass (reduce_warp.cuh): how many run heads (entries that are not a repeat of an earlier entry of
// their row) and row heads (first entry of a row) end up in every 256-entry tile of the SORTED array.  The in-row sort knows
// both for every entry it places; with the counts scanned, a reduce tile knows its place in the output without waiting for
// its predecessors (no look-back chain).  The 32 entries a warp places in one step lie within 32 + 2 * (SEG_MAX - 1)
// positions; a warp of these kernels takes 256 consecutive entries (one tile), so its destinations lie in that tile and its
// two neighbours and the counts are summed in registers.  tile_cnt: head count | row-head count << 31, zeroed by the host.
constexpr int SG_CNT_SHIFT = 8;   // log2 of the reduce pass's tile (RW_TILE)
__global__ void __launch_bounds__(SG_THREADS, 5) k_segment_sort(const u64 *__restrict__ keys_in, const double *__restrict__ vals_in,
                                                                const u32 *n_ptr, int bits_lo, u64 *keys_out, double *vals_out,
                                                                unsigned char *flags, u32 *long_count, u64 *tile_cnt = nullptr,
                                                                int keep_all = 0) {
    // window = the tile plus SEG_MAX entries either side: a row of at most SEG_MAX entries that owns an entry of the
    // tile lies inside it completely
    constexpr int W = SG_TILE + 2 * SEG_MAX;
    constexpr int ITS = (W + SG_THREADS - 1) / SG_THREADS;
    constexpr int NG = ITS * SG_THREADS / 32;  // groups of 32 window entries (the last ones are padding)
    __shared__ u64 s_key[W];
    __shared__ u32 s_col[W];          // column part of every key
    __shared__ u32 s_hb[NG];          // per group: bit l set = entry l starts a row
    __shared__ u32 s_last[NG];        // position of the last row start at or before the end of the group (0: none but entry 0)
    __shared__ u32 s_first[NG];       // position of the first row start inside the group (or beyond the window)
    const u32 n = *n_ptr;
    const u64 base = (u64)blockIdx.x * SG_TILE;
    if (base >= n) return;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (u32 q = tid; q < (u32)W; q += SG_THREADS) {
        const i64 g = (i64)base + (i64)q - SEG_MAX;
        s_key[q] = (g >= 0 && g < (i64)n) ? ld_stream_u64(keys_in + g) : ~0ull;  // ~0: belongs to no row
    }
    // the values are only moved: pull their lines into L2 now, load them when the destinations are known
#pragma unroll
    for (int k = 0; k < SG_IPT; ++k) {
        const u64 g = base + (u64)warp * 256 + (u64)k * 32 + lane;
        if (g < n && (lane & 15) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(vals_in + g));
    }
    __syncthreads();
    const u64 lo_mask = (1ull << bits_lo) - 1;
    // ---- where the rows start: one ballot per group of 32 window entries ------------------------------------------------
#pragma unroll
    for (int it = 0; it < ITS; ++it) {
        const u32 q = (u32)it * SG_THREADS + tid;
        bool head = true;  // padding beyond the window: a row start, so that the last real row ends at the window's edge
        if (q < (u32)W) {
            const u64 kk = s_key[q];
            s_col[q] = (u32)(kk & lo_mask);
            head = q == 0 || (s_key[q - 1] >> bits_lo) != (kk >> bits_lo);
        }
        const u32 hb = __ballot_sync(SPB_FULL_MASK, head);
        if (lane == 0) {
            const u32 g = q >> 5;
            s_hb[g] = hb;
            s_last[g] = hb ? (g << 5) + 31 - __clz(hb) : 0;
            s_first[g] = hb ? (g << 5) + __ffs(hb) - 1 : 0xffffffffu;
        }
    }
    __syncthreads();
    if (warp == 0) {
        // s_last[g] := last row start BEFORE group g; s_first[g] := first row start AFTER group g
        u32 carry = 0;
        for (int r = 0; r < (NG + 31) / 32; ++r) {
            const int g = r * 32 + (int)lane;
            const u32 mine = g < NG ? s_last[g] : 0;
            u32 inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t = __shfl_up_sync(SPB_FULL_MASK, inc, o);
                if (lane >= (u32)o) inc = max(inc, t);
            }
            u32 exc = __shfl_up_sync(SPB_FULL_MASK, inc, 1);
            if (lane == 0) exc = 0;
            if (g < NG) s_last[g] = max(exc, carry);
            carry = max(carry, __shfl_sync(SPB_FULL_MASK, inc, 31));
        }
        carry = 0xffffffffu;
        for (int r = (NG + 31) / 32 - 1; r >= 0; --r) {
            const int g = r * 32 + (int)lane;
            const u32 mine = g < NG ? s_first[g] : 0xffffffffu;
            u32 inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t = __shfl_down_sync(SPB_FULL_MASK, inc, o);
                if (lane + o < 32) inc = min(inc, t);
            }
            u32 exc = __shfl_down_sync(SPB_FULL_MASK, inc, 1);
            if (lane == 31) exc = 0xffffffffu;
            if (g < NG) s_first[g] = min(exc, carry);
            carry = min(carry, __shfl_sync(SPB_FULL_MASK, inc, 0));
        }
    }
    __syncthreads();
    // ---- my entries: position inside the row = entries that must precede it --------------------------------------------
    // (a warp takes 256 CONSECUTIVE entries, 32 per step: all its destinations lie in three tiles of the reduce pass, so the
    // head counts are kept in registers and added once per warp)
    u32 n_long = 0;
    i32 shift_by[SG_IPT];  // destination - source position (inside a row: small); INT32_MIN: entry of a long row
    u32 cnt_at = 0, cnt_before = 0, cnt_after = 0;   // run heads | row heads << 16, by destination tile: mine, the one before, the one after
    const u64 wbase = base + (u64)warp * 256;
#pragma unroll
    for (int k = 0; k < SG_IPT; ++k) {
        const u32 q = SEG_MAX + warp * 256 + (u32)k * 32 + lane;
        const u64 g = wbase + (u64)k * 32 + lane;
        shift_by[k] = 0;
        if (g < n) {
            bool head = true, rhead = false;
            const u32 grp = q >> 5, l = q & 31, hb = s_hb[grp];
            const u32 below = hb & (0xffffffffu >> (31 - l));
            const u32 st = below ? (grp << 5) + 31 - __clz(below) : s_last[grp];
            const u32 above = l < 31 ? hb & (0xffffffffu << (l + 1)) : 0u;
            const u32 nx = above ? (grp << 5) + __ffs(above) - 1 : s_first[grp];   // start of the next row
            const bool is_long = nx - st > (u32)SEG_MAX;  // (a row cut by the window's edge shows more than SEG_MAX entries too)
            if (!is_long) {
                const u32 col = s_col[q];
                u32 before = 0, same = 0;
                for (u32 j = st; j < q; ++j) { before += s_col[j] <= col; same += s_col[j] == col; }   // earlier entries: stable
                for (u32 j = q + 1; j < nx; ++j) before += s_col[j] < col;   // later entries
                shift_by[k] = (i32)(st + before) - (i32)q;
                head = keep_all || same == 0;   // a repeat of an earlier entry of the row is folded into it by the reduce pass
                rhead = before == 0;            // first of its row in column order
            }
            if (flags) flags[g] = is_long;
            n_long += is_long;
            const u64 dst = (u64)((i64)g + shift_by[k]);
            const u32 inc = (head ? 1u : 0u) | (rhead ? 0x10000u : 0u);
            if (dst < wbase) cnt_before += inc;
            else if (dst >= wbase + 256) cnt_after += inc;
            else cnt_at += inc;
        }
    }
    if (tile_cnt) {
        const u32 a0 = __reduce_add_sync(SPB_FULL_MASK, cnt_at), a1 = __reduce_add_sync(SPB_FULL_MASK, cnt_before),
                  a2 = __reduce_add_sync(SPB_FULL_MASK, cnt_after);
        if (lane == 0) {
            u64 *t = tile_cnt + (wbase >> SG_CNT_SHIFT);
            if (a0) atomicAdd((ull *)t, (ull)(a0 & 0xffffu) | ((ull)(a0 >> 16) << 31));
            if (a1) atomicAdd((ull *)(t - 1), (ull)(a1 & 0xffffu) | ((ull)(a1 >> 16) << 31));
            if (a2) atomicAdd((ull *)(t + 1), (ull)(a2 & 0xffffu) | ((ull)(a2 >> 16) << 31));
        }
    }
    double v[SG_IPT];
#pragma unroll
    for (int k = 0; k < SG_IPT; ++k) {
        const u64 g = wbase + (u64)k * 32 + lane;
        v[k] = g < n ? ld_stream_f64(vals_in + g) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < SG_IPT; ++k) {
        const u64 g = wbase + (u64)k * 32 + lane;
        if (g < n) {
            const u64 dst = (u64)((i64)g + shift_by[k]);
            keys_out[dst] = s_key[SEG_MAX + warp * 256 + k * 32 + lane];
            vals_out[dst] = v[k];
        }
    }
    n_long = __reduce_add_sync(SPB_FULL_MASK, n_long);
    if (n_long && lane == 0) atomicAdd(long_count, n_long);
}

// The same for rows of a handful of entries (banded, regridding matrices): every entry simply walks its neighbours in
// the window -- no row table, one barrier.  Measured on the 5-entries-per-row config 5 block: 3.3 ms against 5.0 ms for the
// kernel above; at config 2's 12 entries per row the walk costs 4.5 ms against 2.5 ms.
// (Tried in round 2 and dropped: row bounds from per-group words of row-start bits instead of walking to the row's ends --
// bit-identical, but 23.3 against 21.5 ms per consolidate of the config 5 block, with or without all window loads in flight.)
// A warp takes 256 CONSECUTIVE entries, 32 per step: all its destinations lie in three tiles of the reduce pass, so the head
// counts for that pass are kept in registers and added once per warp.
static_assert(SG_TILE % 256 == 0 && (1 << SG_CNT_SHIFT) == 256, "a warp's 256 consecutive entries are one tile of the reduce pass");
#ifndef SGW_UNROLL_N
#define SGW_UNROLL_N 8
#endif
constexpr int SGW_UNROLL = SGW_UNROLL_N;
__global__ void __launch_bounds__(SG_THREADS, 6) k_segment_sort_walk(const u64 *__restrict__ keys_in, const double *__restrict__ vals_in,
                                                             const u32 *n_ptr, int bits_lo, u64 *keys_out, double *vals_out,
                                                             unsigned char *flags, u32 *long_count, u64 *tile_cnt = nullptr,
                                                             int keep_all = 0) {
    // the window's keys split into their two 32-bit halves (indices are int32): the walk compares rows and columns with
    // 32-bit instructions -- a 64-bit shift + compare per neighbour made this kernel issue-bound
    __shared__ u32 s_row[SG_TILE + 2 * SEG_MAX];
    __shared__ u32 s_col[SG_TILE + 2 * SEG_MAX];
    const u32 n = *n_ptr;
    const u64 base = (u64)blockIdx.x * SG_TILE;
    if (base >= n) return;
    const u32 tid = threadIdx.x;
    const u32 lo_mask = bits_lo >= 32 ? 0xffffffffu : (1u << bits_lo) - 1u;
    for (u32 q = tid; q < SG_TILE + 2 * SEG_MAX; q += SG_THREADS) {
        const i64 g = (i64)base + (i64)q - SEG_MAX;
        u32 row = 0xffffffffu, col = 0;   // row 0xffffffff: belongs to no row (indices are non-negative int32)
        if (g >= 0 && g < (i64)n) {
            const u64 key = ld_stream_u64(keys_in + g);
            row = (u32)(key >> bits_lo);
            col = (u32)key & lo_mask;
        }
        s_row[q] = row;
        s_col[q] = col;
    }
    // a warp takes 256 CONSECUTIVE entries, 32 per step: all its destinations lie in three tiles of the reduce pass, so the
    // head counts are kept in registers and added once per warp
    const u32 lane = tid & 31, warp = tid >> 5;
    const u64 wbase = base + (u64)warp * 256;   // first entry of the warp = first entry of "my" tile
    double v[SG_IPT];
#pragma unroll
    for (int k = 0; k < SG_IPT; ++k) {
        const u64 g = wbase + (u64)k * 32 + lane;
        v[k] = g < n ? ld_stream_f64(vals_in + g) : 0.0;
    }
    __syncthreads();
    u32 n_long = 0;
    u32 cnt_at = 0, cnt_before = 0, cnt_after = 0;   // run heads | row heads << 16, by destination tile: mine, the one before, the one after
#pragma unroll
    for (int k = 0; k < SG_IPT; ++k) {
        const u32 q = SEG_MAX + warp * 256 + (u32)k * 32 + lane;
        const u64 g = wbase + (u64)k * 32 + lane;
        u64 dst = g;
        bool head = true, rhead = false;
        if (g < n) {
            const u32 row = s_row[q], col = s_col[q];
            u32 b = 0, f = 0, before = 0, same = 0;
            // inside a row of more than 2 * SEG_MAX entries: no need to walk to find that out
            if (s_row[q - SEG_MAX] == row || s_row[q + SEG_MAX] == row) b = f = (u32)SEG_MAX;
            else {
                // (unrolled: the neighbours' addresses become immediates, no loop counter in the way -- rows are a handful long)
#pragma unroll SGW_UNROLL
                for (b = 0; b < (u32)SEG_MAX; ++b) {   // earlier entries of my row
                    if (s_row[q - 1 - b] != row) break;
                    const u32 c = s_col[q - 1 - b];
                    before += c <= col;
                    same += c == col;
                }
#pragma unroll SGW_UNROLL
                for (f = 0; f < (u32)SEG_MAX; ++f) {   // later entries of my row
                    if (s_row[q + 1 + f] != row) break;
                    before += s_col[q + 1 + f] < col;
                }
            }
            const bool is_long = b == (u32)SEG_MAX || f == (u32)SEG_MAX || b + f + 1 > (u32)SEG_MAX;
            if (!is_long) {
                dst = g - b + before;
                head = keep_all || same == 0;   // a repeat of an earlier entry of the row is folded into it by the reduce pass
                rhead = before == 0;            // first of its row in column order
            }
            keys_out[dst] = ((u64)row << bits_lo) | col;
            vals_out[dst] = v[k];
            if (flags) flags[g] = is_long;
            n_long += is_long;
            const u32 inc = (head ? 1u : 0u) | (rhead ? 0x10000u : 0u);
            if (dst < wbase) cnt_before += inc;
            else if (dst >= wbase + 256) cnt_after += inc;
            else cnt_at += inc;
        }
    }
    if (tile_cnt) {
        const u32 a0 = __reduce_add_sync(SPB_FULL_MASK, cnt_at), a1 = __reduce_add_sync(SPB_FULL_MASK, cnt_before),
                  a2 = __reduce_add_sync(SPB_FULL_MASK, cnt_after);
        if (lane == 0) {
            u64 *t = tile_cnt + (wbase >> SG_CNT_SHIFT);
            if (a0) atomicAdd((ull *)t, (ull)(a0 & 0xffffu) | ((ull)(a0 >> 16) << 31));
            if (a1) atomicAdd((ull *)(t - 1), (ull)(a1 & 0xffffu) | ((ull)(a1 >> 16) << 31));
            if (a2) atomicAdd((ull *)(t + 1), (ull)(a2 & 0xffffu) | ((ull)(a2 >> 16) << 31));
        }
    }
    n_long = __reduce_add_sync(SPB_FULL_MASK, n_long);
    if (n_long && lane == 0) atomicAdd(long_count, n_long);
}

// The walk for rows of AT MOST SGS_ROW entries (stencil / banded / regridding matrices: 4-5 per row), in registers: a warp step
// looks at 32 consecutive window entries, one per lane; an entry's row mates are at most SGS_ROW - 1 lanes away, so every
// comparison is a shuffle -- no loop, no branch, no shared-memory traffic beyond the entry itself: 2 * (SGS_ROW - 1) row and
// column shuffles per step instead of ~10 divergent loop trips of ~12 instructions.  The first and last SGS_ROW - 1 lanes of a
// step only lend their entries (their own row mates may lie outside the 32), so a step places 32 - 2 * (SGS_ROW - 1) entries;
// a warp takes 256 consecutive entries in 11 such steps.  Whether a step qualifies is decided from one ballot of row-start
// flags: no SGS_ROW consecutive non-starts among the 32 lanes and the entry after them; otherwise (longer rows nearby) that step
// walks through shared memory exactly as k_segment_sort_walk does.  Same outputs, same by-products (long-row count, flags,
// head counts per tile of the reduce pass).
// MEASURED (round 2): bit-identical, but 22.4 ms per consolidate of the config 5 block against 21.1 ms for k_segment_sort_walk --
// an SM has ONE shuffle unit (a warp-wide SHFL per cycle, a quarter of the ALU rate), and 16 shuffles per 24 placed entries
// cost more than the loop trips they replace.  Kept behind SPB_SEGMENT_WALK=2 with its parity cases.
constexpr int SGS_ROW = 5;
constexpr int SGS_H = SGS_ROW - 1;            // lanes either side that only lend their entries
constexpr int SGS_STEP = 32 - 2 * SGS_H;      // entries placed per step
constexpr int SGS_STEPS = (256 + SGS_STEP - 1) / SGS_STEP;
__global__ void __launch_bounds__(SG_THREADS, 6) k_segment_sort_shfl(const u64 *__restrict__ keys_in, const double *__restrict__ vals_in,
                                                                    const u32 *n_ptr, int bits_lo, u64 *keys_out, double *vals_out,
                                                                    unsigned char *flags, u32 *long_count, u64 *tile_cnt = nullptr,
                                                                    int keep_all = 0) {
    __shared__ u32 s_row[SG_TILE + 2 * SEG_MAX];
    __shared__ u32 s_col[SG_TILE + 2 * SEG_MAX];
    __shared__ double s_val[SG_TILE];
    const u32 n = *n_ptr;
    const u64 base = (u64)blockIdx.x * SG_TILE;
    if (base >= n) return;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 lo_mask = bits_lo >= 32 ? 0xffffffffu : (1u << bits_lo) - 1u;
    for (u32 q = tid; q < SG_TILE + 2 * SEG_MAX; q += SG_THREADS) {
        const i64 g = (i64)base + (i64)q - SEG_MAX;
        u32 row = 0xffffffffu, col = 0;   // row 0xffffffff: belongs to no row (indices are non-negative int32)
        if (g >= 0 && g < (i64)n) {
            const u64 key = ld_stream_u64(keys_in + g);
            row = (u32)(key >> bits_lo);
            col = (u32)key & lo_mask;
        }
        s_row[q] = row;
        s_col[q] = col;
    }
#pragma unroll
    for (int k = 0; k < SG_IPT; ++k) {
        const u32 e = (u32)k * SG_THREADS + tid;
        const u64 g = base + e;
        s_val[e] = g < n ? ld_stream_f64(vals_in + g) : 0.0;
    }
    __syncthreads();
    const u64 wbase = base + (u64)warp * 256;   // first entry of the warp = first entry of "my" tile of the reduce pass
    u32 n_long = 0;
    u32 cnt_at = 0, cnt_before = 0, cnt_after = 0;   // run heads | row heads << 16, by destination tile: mine, the one before, the one after
    for (int t = 0; t < SGS_STEPS; ++t) {
        const i32 e = t * SGS_STEP + (i32)lane - SGS_H;          // my entry, relative to the warp's first (halo lanes: outside [0, 256))
        const u32 q = (u32)(SEG_MAX + (i32)warp * 256 + e);      // its window position (>= SEG_MAX - SGS_H)
        const u64 g = wbase + (u64)(i64)e;
        const bool mine = lane >= (u32)SGS_H && lane < (u32)(32 - SGS_H) && e < 256 && g < n;
        const u32 row = s_row[q], col = s_col[q];
        // row starts among my 32 entries and the one after them
        const u32 hb = __ballot_sync(SPB_FULL_MASK, s_row[q - 1] != row);
        const u32 tail = __shfl_sync(SPB_FULL_MASK, (u32)(s_row[q + 1] != row), 31);
        const u64 z = ~((u64)hb | ((u64)tail << 32)) & 0x1ffffffffull;   // non-starts
        u64 run = z;
#pragma unroll
        for (int d = 1; d < SGS_ROW; ++d) run &= z >> d;
        u32 b = 0, f = 0, before = 0, same = 0;
        if (run == 0) {
            // every row that touches a placing lane lies inside the 32: shuffles
#pragma unroll
            for (int d = 1; d < SGS_ROW; ++d) {
                const u32 ru = __shfl_up_sync(SPB_FULL_MASK, row, d), cu = __shfl_up_sync(SPB_FULL_MASK, col, d);
                const u32 rd = __shfl_down_sync(SPB_FULL_MASK, row, d), cd = __shfl_down_sync(SPB_FULL_MASK, col, d);
                const bool up = lane >= (u32)d && ru == row, dn = lane + d < 32u && rd == row;
                b += up;
                before += (up && cu <= col) + (dn && cd < col);
                same += up && cu == col;
                f += dn;
            }
        } else if (mine) {
            // a longer row nearby: walk through shared memory (k_segment_sort_walk)
            if (s_row[q - SEG_MAX] == row || s_row[q + SEG_MAX] == row) b = f = (u32)SEG_MAX;
            else {
                for (b = 0; b < (u32)SEG_MAX; ++b) {
                    if (s_row[q - 1 - b] != row) break;
                    const u32 c = s_col[q - 1 - b];
                    before += c <= col;
                    same += c == col;
                }
                for (f = 0; f < (u32)SEG_MAX; ++f) {
                    if (s_row[q + 1 + f] != row) break;
                    before += s_col[q + 1 + f] < col;
                }
            }
        }
        if (mine) {
            const bool is_long = b == (u32)SEG_MAX || f == (u32)SEG_MAX || b + f + 1 > (u32)SEG_MAX;
            u64 dst = g;
            bool head = true, rhead = false;
            if (!is_long) {
                dst = g - b + before;
                head = keep_all || same == 0;   // a repeat of an earlier entry of the row is folded into it by the reduce pass
                rhead = before == 0;            // first of its row in column order
            }
            keys_out[dst] = ((u64)row << bits_lo) | col;
            vals_out[dst] = s_val[warp * 256 + (u32)e];
            if (flags) flags[g] = is_long;
            n_long += is_long;
            const u32 inc = (head ? 1u : 0u) | (rhead ? 0x10000u : 0u);
            if (dst < wbase) cnt_before += inc;
            else if (dst >= wbase + 256) cnt_after += inc;
            else cnt_at += inc;
        }
    }
    if (tile_cnt) {
        const u32 a0 = __reduce_add_sync(SPB_FULL_MASK, cnt_at), a1 = __reduce_add_sync(SPB_FULL_MASK, cnt_before),
                  a2 = __reduce_add_sync(SPB_FULL_MASK, cnt_after);
        if (lane == 0) {
            u64 *t = tile_cnt + (wbase >> SG_CNT_SHIFT);
            if (a0) atomicAdd((ull *)t, (ull)(a0 & 0xffffu) | ((ull)(a0 >> 16) << 31));
            if (a1) atomicAdd((ull *)(t - 1), (ull)(a1 & 0xffffu) | ((ull)(a1 >> 16) << 31));
            if (a2) atomicAdd((ull *)(t + 1), (ull)(a2 & 0xffffu) | ((ull)(a2 >> 16) << 31));
        }
    }
    n_long = __reduce_add_sync(SPB_FULL_MASK, n_long);
    if (n_long && lane == 0) atomicAdd(long_count, n_long);
}

// entries of the long rows: out of the array (in order) and back
__global__ void k_gather_flagged(const u64 *__restrict__ keys, const double *__restrict__ vals, const unsigned char *__restrict__ flags,
                                 const u64 *__restrict__ slot, u32 n, u64 *k_out, double *v_out) {
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x)
        if (flags[e]) { k_out[slot[e]] = keys[e]; v_out[slot[e]] = vals[e]; }
}
__global__ void k_scatter_flagged(const u64 *__restrict__ k_sorted, const double *__restrict__ v_sorted,
                                  const unsigned char *__restrict__ flags, const u64 *__restrict__ slot, u32 n, u64 *keys, double *vals) {
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x)
        if (flags[e]) { keys[e] = k_sorted[slot[e]]; vals[e] = v_sorted[slot[e]]; }
}
