// gen.cuh -- synthetic inputs generated on the device.  Counter-based:
// every entry is a pure function of (seed, entry number), so the CPU copies in
// oracle/spsparse_oracle.c and spsparse_b200/gen.py produce identical arrays.
#pragma once
#include "common.cuh"

#define GEN_SCRAMBLE 2654435761ull  // Knuth's multiplicative constant (prime): e -> e*c mod n is a
                                    // permutation whenever gcd(c, n) == 1

// config 2 family
__global__ void k_gen_dup_coo(u64 seed, u64 i0, u64 n, u64 ubase, int bits, u64 zero_every, i32 *row,
                              i32 *col, double *val) {
    const u64 mask = (1ull << bits) - 1;
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (u64)gridDim.x * blockDim.x) {
        u64 i = i0 + t;
        u64 s = i < ubase ? i : mix64((seed ^ 0xD0B1Eull) + i) % ubase;
        row[t] = (i32)(mix64(seed + 2 * s) & mask);
        col[t] = (i32)(mix64(seed + 2 * s + 1) & mask);
        double v = 0.5 + u01((seed ^ 0xA11CEull) + i);
        if (zero_every && mix64((seed ^ 0x2E80ull) + i) % zero_every == 0) v = 0.0;
        val[t] = v;
    }
}

// config 5 family: rows [r0, r0 + n/5) of an m x m pentadiagonal matrix.  Logical entry e = 5*(i-r0)+(d+2)
// is stored at slot (e * GEN_SCRAMBLE) mod n.  A diagonal that falls outside the matrix becomes an
// explicit 0.0 on the clamped column -- consolidate() must drop it (algorithm.hpp:284-292).
__global__ void k_gen_banded(u64 seed, u64 m, u64 r0, u64 n, i32 *row, i32 *col, double *val) {
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) {
        u64 i = r0 + e / 5;
        i64 d = (i64)(e % 5) - 2;
        i64 c = (i64)i + d;
        double v = 0.5 + u01(seed + 5 * i + (u64)(d + 2));
        if (c < 0) { c = 0; v = 0.0; }
        if (c >= (i64)m) { c = (i64)m - 1; v = 0.0; }
        u64 slot = (u64)(((unsigned __int128)e * GEN_SCRAMBLE) % n);
        row[slot] = (i32)i;
        col[slot] = (i32)c;
        val[slot] = v;
    }
}

// config 3 family: fine grid ny x nx (rows), coarse grid gy x gx (cols); row r touches the 2x2 block
// of coarse cells at (y*gy/ny, x*gx/nx), clamped at the far edge (which creates duplicate tuples).
__global__ void k_gen_regrid(u64 seed, u32 ny, u32 nx, u32 gy, u32 gx, u64 n, i32 *row, i32 *col, double *val) {
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) {
        u64 r = e / 4;
        u32 q = (u32)(e % 4);
        u32 y = (u32)(r / nx), x = (u32)(r % nx);
        u32 cy = (u32)(((u64)y * gy) / ny), cx = (u32)(((u64)x * gx) / nx);
        u32 yy = cy + (q >> 1), xx = cx + (q & 1);
        if (yy > gy - 1) yy = gy - 1;
        if (xx > gx - 1) xx = gx - 1;
        u64 slot = (u64)(((unsigned __int128)e * GEN_SCRAMBLE) % n);
        row[slot] = (i32)r;
        col[slot] = (i32)((u64)yy * gx + xx);
        val[slot] = 0.1 + 0.9 * u01(seed + e);
    }
}

// config 4 family: R-MAT (a,b,c,d) = (0.57,0.19,0.19,0.05)
__global__ void k_gen_rmat(u64 seed, int scale, u64 n, i32 *row, i32 *col, double *val) {
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) {
        u32 r = 0, c = 0;
        for (int l = 0; l < scale; ++l) {
            double u = u01(seed + (u64)scale * e + (u64)l);
            u32 rb = (u >= 0.76) ? 1u : 0u;                                // c or d quadrant
            u32 cb = ((u >= 0.57 && u < 0.76) || u >= 0.95) ? 1u : 0u;     // b or d quadrant
            r = (r << 1) | rb;
            c = (c << 1) | cb;
        }
        row[e] = (i32)r;
        col[e] = (i32)c;
        val[e] = 0.5 + u01((seed ^ 0x4A77ull) + e);
    }
}

__global__ void k_gen_vector(u64 seed, u64 dim, i32 *idx, double *val) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < dim; j += (u64)gridDim.x * blockDim.x) {
        idx[j] = (i32)j;
        val[j] = 0.5 + u01(seed + j);
    }
}

// payload for sorted_permutation: the entry number itself, carried through the sort as 64 opaque bits
__global__ void k_iota_bits(u64 n, double *out) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        out[i] = __longlong_as_double((long long)i);
}
