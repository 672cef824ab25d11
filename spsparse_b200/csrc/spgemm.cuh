// spgemm.cuh -- row-wise sparse*sparse multiply  C = c * diag(si) * op(A) * diag(sj) * op(B) * diag(sk).
//
// Replaces the loop nest of spsparse::multiply (reference slib/spsparse/multiply_sparse.hpp:192-246)
// and its Join2Xiter/Join3Xiter merge-joins (xiter.hpp:149-278, next_noincr_body.hpp:1-53).  The
// reference evaluates every (row of A) x (column of B) pair; here each non-empty row i of op(A)
// merges the rows B[j,:] of the inner indices j it holds (Gustavson).  Contract kept (SURVEY.md
// App. A M6-M10): rows/cols absent from or zero in scalei/scalek are excluded; j absent from
// scalej is excluded; every output (i,k) adds its terms (a*s)*b in ascending j starting from 0;
// it is emitted iff the sum != 0, as ((sum*C)*a_scale)*b_scale, in (i asc, k asc) order.
//
// Rows are binned by intermediate-product count:
//   short rows (<= MERGE_MAX_LISTS inner indices, <= merge_max_products products): one THREAD per
//     row runs a k-way merge of the (already column-sorted) B rows entirely in registers -- the
//     heads of the lists are compared, equal columns are summed in list (= ascending j) order.
//     Deterministic and bit-identical to the reference's sums.  A symbolic pass counts, a scan
//     places the rows, the numeric pass repeats the merge and writes.
//   long rows: expand-sort-compress through global memory with the radix sort / duplicate-reduce
//     kernels of consolidate (keys (row number, k), stable => terms still in ascending j).
#pragma once
#include "common.cuh"

constexpr int MERGE_MAX_LISTS = 8;
enum { ROW_SKIP = 0, ROW_MERGE = 1, ROW_ESC = 2 };

struct MMOperands {
    // op(A), consolidated, sorted by (row, inner): compressed rows
    const i32 *a_j;
    const double *a_val;
    u32 nnz_a;
    const i32 *arow_id;     // [nrows] row index i of each non-empty row
    const u32 *arow_start;  // [nrows+1]
    u32 nrows;
    // op(B), consolidated, sorted by (inner, col): dense pointer over the inner index
    const u32 *bptr;  // [nj+1]
    const i32 *b_k;
    const double *b_val;
    // dense-ified scale vectors (nullptr when the argument was NULL)
    const double *si;             // [rows of op(A)]  absent = 0 => row excluded
    const double *sj;             // [inner]
    const unsigned char *sj_mask; // [inner]          absent => term excluded
    const double *sk;             // [cols of op(B)]  absent = 0 => column excluded
    double C;
    int debug;  // experiments only (SPB_MERGE_DEBUG): 1 = skip the global stores, 2 = plain (non-streaming) stores
};

// ---- per A entry: number of products it forms (0 if its row or its j is excluded) -------------
__global__ void k_entry_products(MMOperands m, const i32 *__restrict__ a_row, u32 *ent_f) {
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < m.nnz_a; e += (u64)gridDim.x * blockDim.x) {
        i32 j = m.a_j[e];
        u32 f = m.bptr[j + 1] - m.bptr[j];
        if (m.sj_mask && !m.sj_mask[j]) f = 0;
        if (m.si && m.si[a_row[e]] == 0.0) f = 0;
        ent_f[e] = f;
    }
}

// ---- long rows only: product count per row from the per-entry prefix sums -------------------------
// stats: [0] F of merged rows, [1] rows merged, [2] rows ESC, [3] F of ESC rows
__global__ void k_esc_row_products(MMOperands m, const u64 *__restrict__ ent_off, const unsigned char *__restrict__ row_cls,
                                   u64 *esc_f, ull *stats) {
    u64 f_esc = 0;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < m.nrows; r += (u64)gridDim.x * blockDim.x) {
        u64 f = 0;
        if (row_cls[r] == ROW_ESC) f = ent_off[m.arow_start[r + 1]] - ent_off[m.arow_start[r]];
        esc_f[r] = f;
        f_esc += f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) f_esc += __shfl_xor_sync(SPB_FULL_MASK, f_esc, o);
    if (lane_id() == 0 && f_esc) atomicAdd(&stats[3], (ull)f_esc);
}

// ---- short rows: k-way merge in registers, one thread per row -----------------------------------
// The heads of up to NL column-sorted B rows live in registers; each step takes the smallest head
// column, sums every list that holds it in list (= ascending j) order and advances those lists.
template <int NL>
struct RowMerge {
    u32 cur[NL], end[NL];
    i32 hk[NL];
    double hv[NL], as[NL];

    __device__ __forceinline__ void init(const MMOperands &m, u32 s, u32 len) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            cur[l] = end[l] = 0;
            hk[l] = INT32_MAX;
            hv[l] = 0.0;
            as[l] = 0.0;
            if ((u32)l < len) {
                i32 j = __ldg(m.a_j + s + l);
                double a = __ldg(m.a_val + s + l);
                bool ok = !m.sj_mask || m.sj_mask[j];
                if (ok) {
                    cur[l] = __ldg(m.bptr + j);
                    end[l] = __ldg(m.bptr + j + 1);
                    as[l] = m.sj ? __dmul_rn(a, __ldg(m.sj + j)) : a;  // (a*s) first, multiply_sparse.hpp:228
                    if (cur[l] < end[l]) {
                        hk[l] = __ldg(m.b_k + cur[l]);
                        hv[l] = __ldg(m.b_val + cur[l]);
                    }
                }
            }
        }
    }

    // Sum of everything init() loaded.  Callers compare it against an impossible value before entering
    // their merge loop: that one real consumer makes the loop-invariant registers (as[], the row scale)
    // "arrived" for ptxas.  Without it every iteration re-waits on their scoreboard slot, which by then is
    // shared with the freshly issued next-head loads -- the step then stalls a full L2 round trip on the
    // loads it only needs in the NEXT step (profiles/r01_merge_numeric_notes.md).
    __device__ __forceinline__ double touch() const {
        double t = 0.0;
#pragma unroll
        for (int l = 0; l < NL; ++l) t += as[l];
        return t;
    }

    // Next output column and its dot product; false when every list is exhausted.
    __device__ __forceinline__ bool next(const MMOperands &m, i32 &k, double &sum) {
        i32 kmin = hk[0];
#pragma unroll
        for (int l = 1; l < NL; ++l) kmin = min(kmin, hk[l]);
        if (kmin == INT32_MAX) return false;
        double acc = 0.0;  // multiply_sparse.hpp:219
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            if (hk[l] == kmin) {
                acc = __dadd_rn(acc, __dmul_rn(as[l], hv[l]));
                ++cur[l];
                if (cur[l] < end[l]) {
                    hk[l] = __ldg(m.b_k + cur[l]);
                    hv[l] = __ldg(m.b_val + cur[l]);
                } else {
                    hk[l] = INT32_MAX;
                }
            }
        }
        k = kmin;
        sum = acc;
        return true;
    }
};

// keep iff sum != 0 (NaN kept, multiply_sparse.hpp:238) and the column's scale is present and non-zero
__device__ __forceinline__ bool keep_output(const MMOperands &m, i32 k, double sum, double &b_scale) {
    b_scale = 1.0;
    bool keep = (sum != 0.0);
    if (m.sk) {
        b_scale = __ldg(m.sk + k);
        keep = keep && (b_scale != 0.0);
    }
    return keep;
}

// symbolic: bin the row and, for a short row, count its outputs exactly
template <int NL>
__device__ __forceinline__ int count_row(const MMOperands &m, u32 s, u32 len, u32 max_products, u32 &count, u64 &f) {
    RowMerge<NL> st;
    st.init(m, s, len);
    f = 0;
#pragma unroll
    for (int l = 0; l < NL; ++l) f += st.end[l] - st.cur[l];
    count = 0;
    if (f == 0) return ROW_SKIP;
    if (f > max_products) return ROW_ESC;
    i32 k;
    double sum, bs;
    if (st.touch() == -1.2345678e300) return ROW_SKIP;  // never true; see RowMerge::touch
    while (st.next(m, k, sum))
        if (keep_output(m, k, sum, bs)) ++count;
    return ROW_MERGE;
}

// stats: [0] F of merged rows, [1] rows merged, [2] rows ESC
__global__ void __launch_bounds__(128) k_merge_count(MMOperands m, u32 max_products, unsigned char *row_cls,
                                                     u32 *row_cnt, ull *stats) {
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    int cls = ROW_SKIP;
    u32 c = 0;
    u64 f = 0;
    if (r < m.nrows) {
        const u32 s = m.arow_start[r], len = m.arow_start[r + 1] - s;
        const bool on = !m.si || m.si[m.arow_id[r]] != 0.0;  // row excluded by scalei (multiply_sparse.hpp:195)
        if (on) {
            if (len > (u32)MERGE_MAX_LISTS) cls = ROW_ESC;   // products counted later, from the per-entry prefix sums
            else if (len <= 2) cls = count_row<2>(m, s, len, max_products, c, f);
            else if (len <= 4) cls = count_row<4>(m, s, len, max_products, c, f);
            else if (len <= 6) cls = count_row<6>(m, s, len, max_products, c, f);
            else cls = count_row<8>(m, s, len, max_products, c, f);
        }
        row_cls[r] = (unsigned char)cls;
        if (cls != ROW_ESC) row_cnt[r] = c;  // ESC rows are filled by the expand-sort-compress stage
    }
    u64 f_merge = (cls == ROW_MERGE) ? f : 0;
    u32 n_merge = (cls == ROW_MERGE), n_esc = (cls == ROW_ESC);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        f_merge += __shfl_xor_sync(SPB_FULL_MASK, f_merge, o);
        n_merge += __shfl_xor_sync(SPB_FULL_MASK, n_merge, o);
        n_esc += __shfl_xor_sync(SPB_FULL_MASK, n_esc, o);
    }
    if (lane_id() == 0) {
        if (f_merge) atomicAdd(&stats[0], (ull)f_merge);
        if (n_merge) atomicAdd(&stats[1], (ull)n_merge);
        if (n_esc) atomicAdd(&stats[2], (ull)n_esc);
    }
}

// numeric: the 32 rows of a warp advance in lock step; outputs are staged per lane in shared memory and
// flushed as contiguous runs (lane l's run is written by 16 lanes at once), so C is written in full
// sectors instead of one 4/8-byte store per thread per step.
constexpr int MR_THREADS = 128;

template <int NL, int STAGE>
__device__ __forceinline__ void merge_rows_warp(const MMOperands &m, bool mine, u32 s, u32 len, i32 irow,
                                                u64 dst, i32 *sk, double *sv, i32 *c_i, i32 *c_k, double *c_v) {
    constexpr int PITCH = STAGE + 1;
    constexpr int RUNS = 32 / STAGE;  // lanes' runs written per flush iteration
    const u32 lane = lane_id();
    RowMerge<NL> st;
    st.init(m, s, mine ? len : 0);
    double a_scale = 1.0;
    if (mine && m.si) a_scale = m.si[irow];
    u32 cnt = 0;
    bool active = mine;
    if (st.touch() + a_scale + (double)dst == -1.2345678e300) active = false;  // never true; see RowMerge::touch
    for (;;) {
        if (active) {
            i32 k;
            double sum, b_scale;
            if (!st.next(m, k, sum)) active = false;
            else if (keep_output(m, k, sum, b_scale)) {
                sk[lane * PITCH + cnt] = k;
                sv[lane * PITCH + cnt] = __dmul_rn(__dmul_rn(__dmul_rn(sum, m.C), a_scale), b_scale);  // :242
                ++cnt;
            }
        }
        const bool any_active = __any_sync(SPB_FULL_MASK, active);
        if (__any_sync(SPB_FULL_MASK, cnt == (u32)STAGE) || !any_active) {
            __syncwarp();
#pragma unroll 4
            for (int it = 0; it < 32 / RUNS; ++it) {
                const int l = RUNS * it + (int)(lane / STAGE);
                const u32 t = lane % STAGE;
                const u32 n_l = __shfl_sync(SPB_FULL_MASK, cnt, l);
                const u64 d_l = __shfl_sync(SPB_FULL_MASK, dst, l);
                const i32 i_l = __shfl_sync(SPB_FULL_MASK, irow, l);
                if (t < n_l && m.debug != 1) {
                    if (m.debug == 2) {
                        c_k[d_l + t] = sk[l * PITCH + t];
                        c_v[d_l + t] = sv[l * PITCH + t];
                        c_i[d_l + t] = i_l;
                    } else {  // streaming stores: C is written once and never re-read here; keep L1/L2 for B
                        __stcs(c_k + d_l + t, sk[l * PITCH + t]);
                        __stcs(c_v + d_l + t, sv[l * PITCH + t]);
                        __stcs(c_i + d_l + t, i_l);
                    }
                }
            }
            dst += cnt;
            cnt = 0;
            __syncwarp();
        }
        if (!any_active) break;
    }
}

template <int STAGE>
__global__ void __launch_bounds__(MR_THREADS) k_merge_numeric(MMOperands m, const unsigned char *__restrict__ row_cls,
                                                              const u64 *__restrict__ c_ptr, i32 *c_i, i32 *c_k,
                                                              double *c_v) {
    __shared__ i32 s_k[MR_THREADS * (STAGE + 1)];
    __shared__ double s_v[MR_THREADS * (STAGE + 1)];
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 warp = threadIdx.x >> 5;
    const bool mine = (r < m.nrows) && (row_cls[r] == ROW_MERGE);
    u32 s = 0, len = 0;
    i32 irow = 0;
    u64 dst = 0;
    if (mine) {
        s = m.arow_start[r];
        len = m.arow_start[r + 1] - s;
        irow = m.arow_id[r];
        dst = c_ptr[r];
    }
    const u32 maxlen = __reduce_max_sync(SPB_FULL_MASK, len);
    i32 *sk = s_k + warp * 32 * (STAGE + 1);
    double *sv = s_v + warp * 32 * (STAGE + 1);
    if (maxlen == 0) return;
    if (maxlen <= 2) merge_rows_warp<2, STAGE>(m, mine, s, len, irow, dst, sk, sv, c_i, c_k, c_v);
    else if (maxlen <= 4) merge_rows_warp<4, STAGE>(m, mine, s, len, irow, dst, sk, sv, c_i, c_k, c_v);
    else if (maxlen <= 6) merge_rows_warp<6, STAGE>(m, mine, s, len, irow, dst, sk, sv, c_i, c_k, c_v);
    else merge_rows_warp<8, STAGE>(m, mine, s, len, irow, dst, sk, sv, c_i, c_k, c_v);
}

// ---- long rows: expand-sort-compress ------------------------------------------------------------
// chunk boundaries: rb[c] = first row whose ESC product offset >= c*chunk; pb[c] = that offset
__global__ void k_esc_chunks(const u64 *__restrict__ esc_off, u32 nrows, u64 chunk, u32 nchunks, u32 *rb, u64 *pb) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nchunks) return;
    u32 lo = 0, hi = nrows;
    if (c == nchunks) lo = nrows;
    else {
        u64 want = (u64)c * chunk;
        while (lo < hi) {
            u32 mid = lo + (hi - lo) / 2;
            if (esc_off[mid] < want) lo = mid + 1; else hi = mid;
        }
    }
    rb[c] = lo;
    pb[c] = esc_off[lo];
}

// product p of the chunk -> (key = (row number << kbits) | k, value (a*s)*b), in (row, j, k) order
__global__ void k_esc_expand(MMOperands m, const u64 *__restrict__ esc_off, const u64 *__restrict__ ent_off,
                             u32 row_lo, u32 row_hi, u64 p_lo, u64 count, int kbits, u64 *keys, double *vals) {
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (u64)gridDim.x * blockDim.x) {
        const u64 p = p_lo + t;
        // row: last r in [row_lo,row_hi) with esc_off[r] <= p
        u32 lo = row_lo, hi = row_hi;
        while (hi - lo > 1) {
            u32 mid = lo + (hi - lo) / 2;
            if (esc_off[mid] <= p) lo = mid; else hi = mid;
        }
        const u32 r = lo;
        const u32 s = m.arow_start[r], e_end = m.arow_start[r + 1];
        const u64 want = ent_off[s] + (p - esc_off[r]);
        u32 el = s, eh = e_end;  // last entry e with ent_off[e] <= want
        while (eh - el > 1) {
            u32 mid = el + (eh - el) / 2;
            if (ent_off[mid] <= want) el = mid; else eh = mid;
        }
        const u32 e = el;
        const i32 j = m.a_j[e];
        const u32 b = m.bptr[j] + (u32)(want - ent_off[e]);
        double as = m.a_val[e];
        if (m.sj) as = __dmul_rn(as, m.sj[j]);
        keys[t] = ((u64)(r - row_lo) << kbits) | (u64)(u32)m.b_k[b];
        vals[t] = __dmul_rn(as, m.b_val[b]);
    }
}

// compressed rows of the chunk's output: first[row] = first entry, cnt[row] = entries
__global__ void k_esc_row_spans(const i32 *__restrict__ t_row, u32 n, u32 *first, u32 *cnt) {
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (u64)gridDim.x * blockDim.x) {
        i32 r = t_row[t];
        if (t == 0 || t_row[t - 1] != r) first[r] = (u32)t;
    }
}
__global__ void k_esc_row_counts(const i32 *__restrict__ t_row, u32 n, const u32 *__restrict__ first, u32 *cnt) {
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (u64)gridDim.x * blockDim.x) {
        i32 r = t_row[t];
        if (t + 1 == n || t_row[t + 1] != r) cnt[r] = (u32)t + 1 - first[r];
    }
}
__global__ void k_esc_copy(const i32 *__restrict__ t_row, const i32 *__restrict__ t_k, const double *__restrict__ t_v,
                           u32 n, const u32 *__restrict__ first, const u64 *__restrict__ c_ptr,
                           const i32 *__restrict__ arow_id, i32 *c_i, i32 *c_k, double *c_v) {
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (u64)gridDim.x * blockDim.x) {
        i32 r = t_row[t];
        u64 dst = c_ptr[r] + ((u32)t - first[r]);
        c_i[dst] = arow_id[r];
        c_k[dst] = t_k[t];
        c_v[dst] = t_v[t];
    }
}

// ---- matrix * vector  (multiply_sparse.hpp:281-365): one thread per row, ascending j -----------
__global__ void k_mv_rows(MMOperands m, const double *__restrict__ v_dense, const unsigned char *__restrict__ v_mask,
                          double *row_val, unsigned char *row_keep) {
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < m.nrows; r += (u64)gridDim.x * blockDim.x) {
        const i32 irow = m.arow_id[r];
        double a_scale = 1.0;
        bool use = true;
        if (m.si) { a_scale = m.si[irow]; use = (a_scale != 0.0); }
        double sum = 0.0;
        if (use) {
            for (u32 e = m.arow_start[r]; e < m.arow_start[r + 1]; ++e) {
                i32 j = m.a_j[e];
                if (!v_mask[j]) continue;
                if (m.sj_mask && !m.sj_mask[j]) continue;
                double as = m.a_val[e];
                if (m.sj) as = __dmul_rn(as, m.sj[j]);
                sum = __dadd_rn(sum, __dmul_rn(as, v_dense[j]));
            }
        }
        bool keep = use && (sum != 0.0);
        row_keep[r] = keep ? 1 : 0;
        row_val[r] = __dmul_rn(__dmul_rn(sum, m.C), a_scale);
    }
}
__global__ void k_mv_emit(u32 nrows, const i32 *__restrict__ arow_id, const double *__restrict__ row_val,
                          const unsigned char *__restrict__ row_keep, const u32 *__restrict__ slot, i32 *out_i,
                          double *out_v) {
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (u64)gridDim.x * blockDim.x)
        if (row_keep[r]) { out_i[slot[r]] = arow_id[r]; out_v[slot[r]] = row_val[r]; }
}
