// radix_sort9.cuh -- the radix pass of radix_sort.cuh with 9-bit digits (512 buckets).
//
// Why: the passes of a consolidate cover either the whole key or, with the in-row column sort, the row part only.
// A 10^8-row matrix has a 27-bit row part: four 8-bit passes (8+8+8+3), but three 9-bit ones -- one read and one
// write of the whole array less (config 5: 4.0 ms of a 28.5 ms consolidate).  The same holds for any part of
// 9, 17-18 or 25-27 bits (or 33-36, 41-45 ...).  The host picks this variant only when it saves a pass
// (spb_api.cu: sort_reduce; experimental, SPB_RADIX9=1).
//
// What changes against k_radix_pass (the organisation -- one-sweep histogram, warp-level ballot multi-split, staging
// in shared memory, decoupled look-back -- is the same):
//  * nine ballots per item instead of eight;
//  * the warp-private digit counters are 16 bits wide (a warp holds 512 items, a tile 4096), so that 8 x 512 of them
//    take the 8 KB that 8 x 256 32-bit ones take and three blocks still fit an SM;
//  * a thread owns the two adjacent digits 2t and 2t+1: it reads both counters of a warp as one 32-bit word, and the
//    two look-back words of a tile as one 64-bit word (both halves are always published together, so they always carry
//    the same flag).
#pragma once
#include "radix_sort.cuh"

constexpr int R9_BITS = 9;
constexpr int R9_RADIX = 1 << R9_BITS;
constexpr int R9_MAX_PASSES = 7;  // 63 key bits
constexpr int R9_LOOKBACK = 4;   // predecessors read per look-back round trip (64-bit words: half of RS_LOOKBACK keeps the registers level)
constexpr size_t R9_SMEM_BYTES = (size_t)RS_TILE * 16 + (size_t)RS_WARPS * R9_RADIX * sizeof(unsigned short) + 64;

// histogram of every 9-bit digit of every pass (the counterpart of k_sort_hist)
__global__ void __launch_bounds__(512) k_sort_hist9(SortInput in, int passes, int shift0, u32 *hist, u32 *counters) {
    __shared__ u32 s_h[R9_MAX_PASSES * R9_RADIX];
    for (int t = threadIdx.x; t < passes * R9_RADIX; t += blockDim.x) s_h[t] = 0;
    __syncthreads();
    u32 kept = 0, oob = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    constexpr int HU = 4;  // entries per thread per trip, all of their loads issued first
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
            u64 key = pack_key(hi[u], lo[u], in.bits_lo) >> shift0;
            for (int p = 0; p < passes; ++p) {
                atomicAdd(&s_h[p * R9_RADIX + (u32)(key & (R9_RADIX - 1))], 1u);
                key >>= R9_BITS;
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < passes * R9_RADIX; t += blockDim.x)
        if (s_h[t]) atomicAdd(&hist[t], s_h[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(SPB_FULL_MASK, kept, o);
    if (lane_id() == 0 && kept) atomicAdd(&counters[0], kept);
    if (oob) counters[1] = 1;
}

// The same histogram with 128-bit loads: a thread takes FOUR CONSECUTIVE entries per trip -- one 16-byte load of row indices,
// one of column indices, two of values -- and two trips' loads are in flight before the first counter is touched.  Needs
// 16-byte aligned arrays (the host falls back to k_sort_hist / k_sort_hist9 otherwise); the last n % 4 entries are taken
// one by one.  BITS = 8 or 9.
template <int BITS>
__global__ void __launch_bounds__(512) k_sort_hist_v4(SortInput in, int passes, int shift0, u32 *hist, u32 *counters) {
    constexpr int RADIX = 1 << BITS;
    constexpr int MAXP = BITS == 9 ? R9_MAX_PASSES : RS_MAX_PASSES;
    __shared__ u32 s_h[MAXP * RADIX];
    for (int t = threadIdx.x; t < passes * RADIX; t += blockDim.x) s_h[t] = 0;
    __syncthreads();
    u32 kept = 0, oob = 0;
    auto count = [&](u64 i, i32 hi, i32 lo, double v) {
        if ((u32)hi >= in.extent_hi || (u32)lo >= in.extent_lo) { oob = 1; return; }
        if (!input_kept(in, (u32)i, v)) return;
        ++kept;
        u64 key = pack_key(hi, lo, in.bits_lo) >> shift0;
        for (int p = 0; p < passes; ++p) {
            atomicAdd(&s_h[p * RADIX + (u32)(key & (RADIX - 1))], 1u);
            key >>= BITS;
        }
    };
    const u64 quads = in.n / 4;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const int4 *hi4 = reinterpret_cast<const int4 *>(in.hi), *lo4 = reinterpret_cast<const int4 *>(in.lo);
    const double2 *v2 = reinterpret_cast<const double2 *>(in.val);
    constexpr int HU = 2;
    for (u64 q0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; q0 < quads; q0 += HU * stride) {
        int4 h[HU], l[HU];
        double2 va[HU], vb[HU];
#pragma unroll
        for (int u = 0; u < HU; ++u) {
            const u64 q = q0 + u * stride < quads ? q0 + u * stride : q0;
            asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(h[u].x), "=r"(h[u].y), "=r"(h[u].z), "=r"(h[u].w) : "l"(hi4 + q));
            if (in.lo) asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(l[u].x), "=r"(l[u].y), "=r"(l[u].z), "=r"(l[u].w) : "l"(lo4 + q));
            else l[u] = make_int4(0, 0, 0, 0);
            asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(va[u].x), "=d"(va[u].y) : "l"(v2 + 2 * q));
            asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(vb[u].x), "=d"(vb[u].y) : "l"(v2 + 2 * q + 1));
        }
#pragma unroll
        for (int u = 0; u < HU; ++u) {
            const u64 q = q0 + u * stride;
            if (q >= quads) break;
            count(4 * q, h[u].x, l[u].x, va[u].x);
            count(4 * q + 1, h[u].y, l[u].y, va[u].y);
            count(4 * q + 2, h[u].z, l[u].z, vb[u].x);
            count(4 * q + 3, h[u].w, l[u].w, vb[u].y);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (in.n & 3)) {
        const u64 i = quads * 4 + threadIdx.x;
        count(i, in.hi[i], in.lo ? in.lo[i] : 0, in.val[i]);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < passes * RADIX; t += blockDim.x)
        if (s_h[t]) atomicAdd(&hist[t], s_h[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(SPB_FULL_MASK, kept, o);
    if (lane_id() == 0 && kept) atomicAdd(&counters[0], kept);
    if (oob) counters[1] = 1;
}

// exclusive scan of each pass's 512 counts -> first output slot of each digit (in place); thread t: digits 2t, 2t+1
__global__ void __launch_bounds__(R9_RADIX / 2) k_bucket_starts9(u32 *hist) {
    __shared__ u32 s_w[R9_RADIX / 64];
    u32 *h = hist + blockIdx.x * R9_RADIX;
    const u32 t = threadIdx.x;
    const u32 c0 = h[2 * t], c1 = h[2 * t + 1];
    const u32 incl = warp_incl_scan(c0 + c1);
    if (lane_id() == 31) s_w[t >> 5] = incl;
    __syncthreads();
    u32 add = 0;
    for (u32 w = 0; w < (t >> 5); ++w) add += s_w[w];
    h[2 * t] = add + incl - c0 - c1;
    h[2 * t + 1] = add + incl - c1;
}

// PassArgs as for k_radix_pass, with bucket_start[512] and lookback[tiles][512] (8-byte aligned); rank_mode unused.
template <bool PASS0, bool BULK = false>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_BLOCKS) k_radix_pass9(PassArgs a, SortInput in) {
    __shared__ __align__(8) u64 s_mbar[2];   // BULK: see k_radix_pass
    extern __shared__ __align__(128) unsigned char smem_raw9[];
    u64 *s_keys = reinterpret_cast<u64 *>(smem_raw9);
    double *s_vals = reinterpret_cast<double *>(s_keys + RS_TILE);
    u32 *s_cnt2 = reinterpret_cast<u32 *>(s_vals + RS_TILE);   // [RS_WARPS][256]: counters of digits 2t (low half) and 2t+1; later: global bases [512]
    unsigned short *s_cnt16 = reinterpret_cast<unsigned short *>(s_cnt2);  // the same memory, one counter per element
    u32 *s_misc = s_cnt2 + RS_WARPS * (R9_RADIX / 2);          // [0] tile, [1..8] warp sums, [12] valid items

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_misc[0] = atomicAdd(a.ticket, 1u);
        if (BULK) { mbar_init(&s_mbar[0], 1); mbar_init(&s_mbar[1], 1); mbar_fence_init(); }
    }
    for (int t = tid; t < RS_WARPS * (R9_RADIX / 2); t += RS_THREADS) s_cnt2[t] = 0;
    __syncthreads();
    const u32 tile = s_misc[0];
    const u32 n = PASS0 ? in.n : *a.n_ptr;
    const u64 tile_base = (u64)tile * RS_TILE;
    if (tile_base >= n) return;
    if (BULK && PASS0 && tid == 0) {   // see k_radix_pass
        i32 *s_idx = reinterpret_cast<i32 *>(s_keys);
        mbar_expect_tx(&s_mbar[0], (in.lo ? 2u : 1u) * RS_TILE * 4u);
        bulk_load(s_idx, in.hi + tile_base, RS_TILE * 4u, &s_mbar[0]);
        if (in.lo) bulk_load(s_idx + RS_TILE, in.lo + tile_base, RS_TILE * 4u, &s_mbar[0]);
        mbar_expect_tx(&s_mbar[1], RS_TILE * 8u);
        bulk_load(s_vals, in.val + tile_base, RS_TILE * 8u, &s_mbar[1]);
    }
    if (BULK && !PASS0 && tid == 0) {
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
#pragma unroll
        for (int k = 0; k < RS_IPT; ++k) {
            u64 i = wbase + (u64)k * 32;
            if (i < n && (lane & 15) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.vals_in + i));
        }
    }

    // ---- rank inside the warp: ballot multi-split over nine digit bits, warp-private 16-bit counters ---------
    unsigned short *mycnt = s_cnt16 + warp * R9_RADIX;
    const u32 lt = lanemask_lt();
    unsigned short pos[RS_IPT];
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        const bool ok = (valid_bits >> k) & 1u;
        const u32 d = (u32)((key[k] >> a.shift) & (R9_RADIX - 1));
        u32 peers = __ballot_sync(SPB_FULL_MASK, ok);
#pragma unroll
        for (int b = 0; b < R9_BITS; ++b) {
            int sgn;  // all-ones if bit b of the digit is set
            asm("bfe.s32 %0, %1, %2, 1;" : "=r"(sgn) : "r"(d), "r"(b));
            const u32 m = __ballot_sync(SPB_FULL_MASK, sgn != 0);
            peers &= ~(m ^ (u32)sgn);
        }
        const u32 leader = ok ? (u32)(__ffs(peers) - 1) : lane;
        u32 before = 0;
        if (ok && lane == leader) {   // one lane per distinct digit: the 16-bit stores of two leaders never touch the same element
            before = mycnt[d];
            mycnt[d] = (unsigned short)(before + __popc(peers));
        }
        before = __shfl_sync(SPB_FULL_MASK, before, leader);
        pos[k] = (unsigned short)(before + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit pair (thread t: digits 2t, 2t+1): tile totals, start inside the tile, offsets of each warp ----
    u32 tot0 = 0, tot1 = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
        const u32 c = s_cnt2[w * (R9_RADIX / 2) + tid];
        tot0 += c & 0xffffu;
        tot1 += c >> 16;
    }
    // publish this tile's counts of both digits as early as possible (one 64-bit word: both halves carry the same flag)
    u64 *lb_all = reinterpret_cast<u64 *>(a.lookback);
    u64 *lb = lb_all + (u64)tile * (R9_RADIX / 2) + tid;
    {
        const u32 w0 = tile == 0 ? RS_WORD_INCL(tot0) : RS_WORD_AGG(tot0), w1 = tile == 0 ? RS_WORD_INCL(tot1) : RS_WORD_AGG(tot1);
        st_relaxed_u64(lb, (u64)w0 | ((u64)w1 << 32));
    }
    const u32 pair = tot0 + tot1;
    const u32 incl = warp_incl_scan(pair);
    if (lane == 31) s_misc[1 + warp] = incl;
    __syncthreads();
    u32 lstart0 = incl - pair;
    for (u32 w = 0; w < warp; ++w) lstart0 += s_misc[1 + w];
    const u32 lstart1 = lstart0 + tot0;
    {
        u32 run0 = lstart0, run1 = lstart1;   // at most 4096: fits the 16-bit halves
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const u32 c = s_cnt2[w * (R9_RADIX / 2) + tid];
            s_cnt2[w * (R9_RADIX / 2) + tid] = (run0 & 0xffffu) | (run1 << 16);
            run0 += c & 0xffffu;
            run1 += c >> 16;
        }
    }
    if (tid == RS_THREADS - 1) s_misc[12] = lstart1 + tot1;
    __syncthreads();
    const u32 nvalid = s_misc[12];

    // ---- stage keys and values in digit order in shared memory --------------------------------
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        if ((valid_bits >> k) & 1u) {
            const u32 d = (u32)((key[k] >> a.shift) & (R9_RADIX - 1));
            const u32 p = (u32)mycnt[d] + pos[k];
            pos[k] = (unsigned short)p;
            s_keys[p] = key[k];
        }
    }
    {
        const double *vsrc = PASS0 ? in.val : a.vals_in;
        double v[RS_IPT];
        if (BULK) {
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

    // ---- decoupled look-back, one digit pair per thread ---------------------------------------------
    u32 excl0 = 0, excl1 = 0;
    if (tile > 0) {
        const u64 both_incl = (u64)RS_WORD_INCL(0u) | ((u64)RS_WORD_INCL(0u) << 32);
        i64 p = (i64)tile - 1;
        bool done = false;
        while (!done) {
            u64 w[R9_LOOKBACK];
#pragma unroll
            for (int u = 0; u < R9_LOOKBACK; ++u)
                w[u] = (p - u >= 0) ? ld_relaxed_u64(lb_all + (u64)(p - u) * (R9_RADIX / 2) + tid) : both_incl;
#pragma unroll
            for (int u = 0; u < R9_LOOKBACK; ++u) {
                if (done) break;
                while (!RS_READY((u32)w[u])) w[u] = ld_relaxed_u64(lb_all + (u64)(p - u) * (R9_RADIX / 2) + tid);
                const u32 w0 = (u32)w[u], w1 = (u32)(w[u] >> 32);
                excl0 += RS_VALUE(w0);
                excl1 += RS_VALUE(w1);
                if (RS_IS_INCL(w0)) done = true;
            }
            p -= R9_LOOKBACK;
        }
        st_relaxed_u64(lb, (u64)RS_WORD_INCL(excl0 + tot0) | ((u64)RS_WORD_INCL(excl1 + tot1) << 32));
    }
    const u32 gbase0 = a.bucket_start[2 * tid] + excl0 - lstart0;  // global slot = gbase + position in tile
    const u32 gbase1 = a.bucket_start[2 * tid + 1] + excl1 - lstart1;
    __syncthreads();  // all staging done, the counters are free
    s_cnt2[2 * tid] = gbase0;
    s_cnt2[2 * tid + 1] = gbase1;
    __syncthreads();

    // ---- write out: consecutive threads -> consecutive staged items -> runs of consecutive slots -
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        u32 p = (u32)k * RS_THREADS + tid;
        if (p < nvalid) {
            u64 kk = s_keys[p];
            u32 dst = s_cnt2[(u32)((kk >> a.shift) & (R9_RADIX - 1))] + p;
            a.keys_out[dst] = kk;
            a.vals_out[dst] = s_vals[p];
        }
    }
}
