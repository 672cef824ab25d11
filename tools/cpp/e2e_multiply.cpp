// e2e_multiply.cpp -- what a user of the reference's C++ API sees end to end: BASELINE config 5 held in VectorCooArrays
// (std::vector storage, pageable host memory), one call of spsparse::multiply per step -- the reference's own entry point
// (slib/spsparse/multiply_sparse.hpp:152-164), here through include/spsparse/.  Every step uploads A, B and w, consolidates,
// multiplies and downloads C into a fresh VectorCooArray.  bench.py runs this and reports it under e2e.cpp_api.
//     e2e_multiply [rows = 100000000] [steps = 2] [warmup = 1]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include <spsparse/VectorCooArray.hpp>
#include <spsparse/multiply_sparse.hpp>

typedef unsigned long long u64;
static inline u64 mix64(u64 x) {   // SURVEY.md Appendix C (splitmix64 finaliser), as in spsparse_b200/gen.py
    u64 z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(u64 x) { return (double)(mix64(x) >> 11) * 1.1102230246251565e-16; }

typedef spsparse::VectorCooArray<int, double, 2> Mat;
typedef spsparse::VectorCooArray<int, double, 1> Vec;

// gen.banded: 5 entries per row, inserted in a scrambled order, explicit zeros on the clamped edge diagonals
static void banded(Mat &M, u64 seed, long m) {
    const u64 n = 5ull * (u64)m;
    int *ip[2];
    double *vp;
    M.set_shape({(size_t)m, (size_t)m});
    M.grow_raw(n, ip, &vp);
    const unsigned T = std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t)
        th.emplace_back([=]() {
            for (u64 e = n * t / T; e < n * (t + 1) / T; ++e) {
                const long i = (long)(e / 5), d = (long)(e % 5) - 2;
                long c = i + d;
                double v = 0.5 + u01(seed + (u64)(5 * i + d + 2));
                if (c < 0 || c >= m) { v = 0.0; c = c < 0 ? 0 : m - 1; }
                const u64 slot = (u64)(((unsigned __int128)e * 2654435761ull) % n);
                ip[0][slot] = (int)i; ip[1][slot] = (int)c; vp[slot] = v;
            }
        });
    for (auto &x : th) x.join();
}

int main(int argc, char **argv) {
    const long m = argc > 1 ? atol(argv[1]) : 100000000L;
    const int steps = argc > 2 ? atoi(argv[2]) : 2, warmup = argc > 3 ? atoi(argv[3]) : 1;
    Mat A, B;
    Vec w;
    banded(A, 0x5EED0005ull, m);
    banded(B, 0x5EED0015ull, m);
    {
        int *ip[1];
        double *vp;
        w.set_shape({(size_t)m});
        w.grow_raw((size_t)m, ip, &vp);
        for (long j = 0; j < m; ++j) { ip[0][j] = (int)j; vp[j] = 0.5 + u01(0x5EED0025ull + (u64)j); }
        w.set_sorted({0});
    }
    double best = 1e300, sum = 0;
    u64 nnz = 0, chk = 0;
    for (int s = 0; s < warmup + steps; ++s) {
        Mat C;
        const auto t0 = std::chrono::steady_clock::now();
        spsparse::multiply(C, 1.0, (Vec *)NULL, A, '.', &w, B, '.', (Vec *)NULL);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (s >= warmup) { best = ms < best ? ms : best; sum += ms; }
        if (s == warmup + steps - 1) {
            nnz = C.size();
            for (size_t t = 0; t < C.size(); ++t) chk += (u64)C.index(0, t) * 1000003ull + (u64)C.index(1, t);
        }
    }
    printf("{\"rows\": %ld, \"steps\": %d, \"ms_per_step\": %.3f, \"ms_best\": %.3f, \"nnz_c\": %llu, \"index_checksum\": %llu, "
           "\"h2d_bytes_per_step\": %.0f, \"d2h_bytes_per_step\": %.0f}\n",
           m, steps, sum / steps, best, nnz, chk, 16.0 * 5 * m * 2 + 12.0 * m, 16.0 * (double)nnz);
    return 0;
}
