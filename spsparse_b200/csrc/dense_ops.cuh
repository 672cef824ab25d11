// dense_ops.cuh -- the streaming steps either side of the hot path:
//   to_dense   VectorCooArray::to_dense (reference slib/spsparse/VectorCooArray.hpp:313-321) = a zeroed row-major
//              array + copy() (algorithm.hpp:30-37) into a DenseAccum (accum.hpp:110-140)
//   to_sparse  algorithm.hpp:433-440: every element != 0 of a dense array, in storage order
// (transpose and copy, algorithm.hpp:30-57, are device-to-device copies of the index vectors: no kernel.)
#pragma once
#include "common.cuh"
#include "reduce_by_key.cuh"

// Entries sorted stably by (i, k) -- duplicates of a cell are adjacent and in insertion order.  The first entry
// of every cell replays DenseAccum::add over its run, starting from the 0 the array was filled with, so that the
// cell ends up bit-identical to the reference's sequential loop (ADD: ((0 + v1) + v2) + ...; REPLACE: the last
// value; LEAVE_ALONE as written in accum.hpp:128-130: overwrite unless the cell holds a NaN).
__global__ void k_dense_fold(const i32 *__restrict__ si, const i32 *__restrict__ sk, const double *__restrict__ sv, u64 n,
                             u64 ncols, int policy, double *dense) {
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (u64)gridDim.x * blockDim.x) {
        const i32 i = si[t], k = sk ? sk[t] : 0;
        if (t && si[t - 1] == i && (!sk || sk[t - 1] == k)) continue;  // not the first entry of its cell
        double cell = 0.0;
        for (u64 u = t; u < n && si[u] == i && (!sk || sk[u] == k); ++u) {
            const double v = sv[u];
            if (policy == POLICY_LEAVE_ALONE) { if (!isnan(cell)) cell = v; }
            else if (policy == POLICY_ADD) cell = __dadd_rn(cell, v);
            else cell = v;
        }
        dense[(u64)i * ncols + (u64)k] = cell;
    }
}

__global__ void k_dense_flags(const double *__restrict__ dense, u64 cells, unsigned char *keep) {
    for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (u64)gridDim.x * blockDim.x)
        keep[c] = dense[c] != 0.0;  // NaN != 0: kept (algorithm.hpp:438)
}

__global__ void k_dense_compact(const double *__restrict__ dense, u64 cells, u64 ncols, const unsigned char *__restrict__ keep,
                                const u64 *__restrict__ slot, i32 *out_i, i32 *out_k, double *out_v) {
    for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (u64)gridDim.x * blockDim.x) {
        if (keep[c]) {
            const u64 d = slot[c];
            out_i[d] = (i32)(c / ncols);
            if (out_k) out_k[d] = (i32)(c % ncols);
            out_v[d] = dense[c];
        }
    }
}
