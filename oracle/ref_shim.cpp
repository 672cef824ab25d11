// ref_shim.cpp -- C-ABI wrapper around the GENUINE reference headers (TEST INFRASTRUCTURE ONLY).
//
// Compiled by oracle/Makefile against /root/reference/slib (sources stay where they lie; nothing
// from the reference is copied into this repo) into oracle/_ref/libspsparse_ref.so.  The entry
// points mirror oracle/spsparse_oracle.c one-for-one so that the tests can run the restatement
// and the real thing on the same inputs, and so bench.py can time the real thing as the CPU
// baseline.  Reference calls made here:
//   spsparse::sorted_permutation  slib/spsparse/algorithm.hpp:411-427
//   spsparse::consolidate         slib/spsparse/algorithm.hpp:251-319
//   spsparse::dim_beginnings      slib/spsparse/algorithm.hpp:74-118
//   spsparse::Join2Xiter/Join3Xiter slib/spsparse/xiter.hpp:149-278
//   spsparse::multiply (MM, MV)   slib/spsparse/multiply_sparse.hpp:152-248, 281-365
//   spsparse::transpose           slib/spsparse/algorithm.hpp:46-57
//   spsparse::copy into DenseAccum (= to_dense with a policy)  algorithm.hpp:30-37, accum.hpp:110-140
//   spsparse::to_sparse           slib/spsparse/algorithm.hpp:433-440
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <spsparse/VectorCooArray.hpp>
#include <spsparse/multiply_sparse.hpp>

using namespace spsparse;
typedef VectorCooArray<int, double, 2> Mat;
typedef VectorCooArray<int, double, 1> Vec;

struct orc_mat {
    uint64_t shape[2];
    int64_t n;
    const int32_t *idx0, *idx1;
    const double *val;
    int sort_order[2];
};
struct orc_vec {
    uint64_t shape;
    int64_t n;
    const int32_t *idx;
    const double *val;
    int sort_order;
};

static DuplicatePolicy pol(int p) {
    return p == 0 ? DuplicatePolicy::LEAVE_ALONE : (p == 1 ? DuplicatePolicy::ADD : DuplicatePolicy::REPLACE);
}

static void fill(Mat &M, const orc_mat *m) {
    M.set_shape({(size_t)m->shape[0], (size_t)m->shape[1]});
    M.reserve((size_t)m->n);
    for (int64_t i = 0; i < m->n; ++i) M.add({m->idx0[i], m->idx1[i]}, m->val[i]);
    if (m->sort_order[0] >= 0) M.set_sorted({m->sort_order[0], m->sort_order[1]});
}
static void fill(Vec &V, const orc_vec *v) {
    V.set_shape({(size_t)v->shape});
    V.reserve((size_t)v->n);
    for (int64_t i = 0; i < v->n; ++i) V.add({v->idx[i]}, v->val[i]);
    if (v->sort_order >= 0) V.set_sorted({v->sort_order});
}

extern "C" {

void ref_free(void *p) { free(p); }

void ref_sorted_permutation(int rank, int64_t n, const int32_t *idx0, const int32_t *idx1,
                            const int *sort_order, int64_t *perm) {
    if (rank == 2) {
        Mat A({(size_t)1 << 31, (size_t)1 << 31});
        for (int64_t i = 0; i < n; ++i) A.add({idx0[i], idx1[i]}, 1.0);
        auto p = sorted_permutation(A, {sort_order[0], sort_order[1]});
        for (int64_t i = 0; i < n; ++i) perm[i] = (int64_t)p[i];
    } else {
        Vec A({(size_t)1 << 31});
        for (int64_t i = 0; i < n; ++i) A.add({idx0[i]}, 1.0);
        auto p = sorted_permutation(A, {sort_order[0]});
        for (int64_t i = 0; i < n; ++i) perm[i] = (int64_t)p[i];
    }
}

int64_t ref_consolidate(int rank, int64_t n, const int32_t *idx0, const int32_t *idx1,
                        const double *val, const int *sort_order, int policy, int zero_nan,
                        int32_t *out0, int32_t *out1, double *outv) {
    if (rank == 2) {
        Mat A({(size_t)1 << 31, (size_t)1 << 31});
        A.reserve((size_t)n);
        for (int64_t i = 0; i < n; ++i) A.add({idx0[i], idx1[i]}, val[i]);
        Mat R(A.shape);
        consolidate(R, A, {sort_order[0], sort_order[1]}, pol(policy), zero_nan != 0);
        for (size_t i = 0; i < R.size(); ++i) { out0[i] = R.index(0, i); out1[i] = R.index(1, i); outv[i] = R.val(i); }
        return (int64_t)R.size();
    } else {
        Vec A({(size_t)1 << 31});
        A.reserve((size_t)n);
        for (int64_t i = 0; i < n; ++i) A.add({idx0[i]}, val[i]);
        Vec R(A.shape);
        consolidate(R, A, {sort_order[0]}, pol(policy), zero_nan != 0);
        for (size_t i = 0; i < R.size(); ++i) { out0[i] = R.index(0, i); outv[i] = R.val(i); }
        return (int64_t)R.size();
    }
}

void ref_transpose(int rank, int64_t n, const int32_t *idx0, const int32_t *idx1, const int *perm,
                   int32_t *out0, int32_t *out1) {
    if (rank == 2) {
        Mat A({(size_t)1 << 31, (size_t)1 << 31});
        for (int64_t i = 0; i < n; ++i) A.add({idx0[i], idx1[i]}, 1.0);
        Mat R(A.shape);
        transpose(R, A, {perm[0], perm[1]});
        for (size_t i = 0; i < R.size(); ++i) { out0[i] = R.index(0, i); out1[i] = R.index(1, i); }
    } else {
        Vec A({(size_t)1 << 31});
        for (int64_t i = 0; i < n; ++i) A.add({idx0[i]}, 1.0);
        Vec R(A.shape);
        transpose(R, A, {perm[0]});
        for (size_t i = 0; i < R.size(); ++i) out0[i] = R.index(0, i);
    }
}

int ref_to_dense(int rank, const uint64_t *shape, int64_t n, const int32_t *idx0, const int32_t *idx1,
                 const double *val, int policy, double *dense) {
    if (rank == 2) {
        Mat A({(size_t)shape[0], (size_t)shape[1]});
        for (int64_t i = 0; i < n; ++i) A.add({idx0[i], idx1[i]}, val[i]);
        blitz::Array<double, 2> ret(ibmisc::to_tiny<int, size_t, 2>(A.shape));
        ret = 0;
        DenseAccum<Mat> accum(ret, pol(policy));
        copy(accum, A);
        for (size_t i = 0; i < shape[0]; ++i)
            for (size_t j = 0; j < shape[1]; ++j) dense[i * shape[1] + j] = ret((int)i, (int)j);
    } else {
        Vec A({(size_t)shape[0]});
        for (int64_t i = 0; i < n; ++i) A.add({idx0[i]}, val[i]);
        blitz::Array<double, 1> ret(ibmisc::to_tiny<int, size_t, 1>(A.shape));
        ret = 0;
        DenseAccum<Vec> accum(ret, pol(policy));
        copy(accum, A);
        for (size_t i = 0; i < shape[0]; ++i) dense[i] = ret((int)i);
    }
    return 0;
}

int64_t ref_to_sparse(int rank, const uint64_t *shape, const double *dense, int32_t *out0, int32_t *out1,
                      double *outv) {
    if (rank == 2) {
        blitz::Array<double, 2> arr((int)shape[0], (int)shape[1]);
        for (size_t i = 0; i < shape[0]; ++i)
            for (size_t j = 0; j < shape[1]; ++j) arr((int)i, (int)j) = dense[i * shape[1] + j];
        Mat R({(size_t)shape[0], (size_t)shape[1]});
        to_sparse(R, arr);
        for (size_t i = 0; i < R.size(); ++i) { out0[i] = R.index(0, i); out1[i] = R.index(1, i); outv[i] = R.val(i); }
        return (int64_t)R.size();
    }
    blitz::Array<double, 1> arr((int)shape[0]);
    for (size_t i = 0; i < shape[0]; ++i) arr((int)i) = dense[i];
    Vec R({(size_t)shape[0]});
    to_sparse(R, arr);
    for (size_t i = 0; i < R.size(); ++i) { out0[i] = R.index(0, i); outv[i] = R.val(i); }
    return (int64_t)R.size();
}

// Times only the reference's consolidate() call (container fill excluded); returns seconds.
double ref_consolidate_timed(int64_t n, const int32_t *idx0, const int32_t *idx1, const double *val,
                             const int *sort_order, int policy, int zero_nan, int64_t *out_n,
                             double *out_sum) {
    Mat A({(size_t)1 << 31, (size_t)1 << 31});
    A.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) A.add({idx0[i], idx1[i]}, val[i]);
    Mat R(A.shape);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    consolidate(R, A, {sort_order[0], sort_order[1]}, pol(policy), zero_nan != 0);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double s = 0;
    for (size_t i = 0; i < R.size(); ++i) s += R.val(i);
    *out_n = (int64_t)R.size();
    *out_sum = s;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

// idx_dim is unused by the genuine call; the sorted array is rebuilt from (idx0, idx1).
int64_t ref_dim_beginnings(int64_t n, const int32_t *idx0, const int32_t *idx1, const int *sort_order,
                           int64_t *out) {
    Mat A({(size_t)1 << 31, (size_t)1 << 31});
    for (int64_t i = 0; i < n; ++i) A.add({idx0[i], idx1[i]}, 1.0);
    A.set_sorted({sort_order[0], sort_order[1]});
    auto db = dim_beginnings(A);
    for (size_t i = 0; i < db.size(); ++i) out[i] = (int64_t)db[i];
    return (int64_t)db.size();
}

int64_t ref_join(const int32_t *a, int64_t na, const int32_t *b, int64_t nb, const int32_t *c,
                 int64_t nc, int32_t *out) {
    typedef STLXiter<std::vector<int>::iterator> X;
    std::vector<int> va(a, a + na), vb(b, b + nb), vc;
    int64_t m = 0;
    if (nc < 0) {
        for (auto ii(Join2Xiter<X, X>(X(va.begin(), va.end()), X(vb.begin(), vb.end()))); !ii.eof(); ++ii)
            out[m++] = *ii.i1;
    } else {
        vc.assign(c, c + nc);
        for (auto ii(Join3Xiter<X, X, X>(X(va.begin(), va.end()), X(vb.begin(), vb.end()),
                                         X(vc.begin(), vc.end())));
             !ii.eof(); ++ii)
            out[m++] = *ii.i1;
    }
    return m;
}

static int export_mat(Mat const &R, uint64_t out_shape[2], int64_t *out_n, int32_t **oi, int32_t **ok,
                      double **ov) {
    out_shape[0] = R.shape[0];
    out_shape[1] = R.shape[1];
    size_t n = R.size();
    *out_n = (int64_t)n;
    *oi = (int32_t *)malloc((n ? n : 1) * sizeof(int32_t));
    *ok = (int32_t *)malloc((n ? n : 1) * sizeof(int32_t));
    *ov = (double *)malloc((n ? n : 1) * sizeof(double));
    for (size_t i = 0; i < n; ++i) { (*oi)[i] = R.index(0, i); (*ok)[i] = R.index(1, i); (*ov)[i] = R.val(i); }
    return 0;
}

// Returns 0 ok, 1 inner-dimension error (spsparse::Exception thrown by the default handler).
int ref_multiply_mm(double C, const orc_vec *si, const orc_mat *A, char tA, const orc_vec *sj,
                    const orc_mat *B, char tB, const orc_vec *sk, int policy, int zero_nan,
                    uint64_t out_shape[2], int64_t *out_n, int32_t **out_i, int32_t **out_k,
                    double **out_v, double *seconds) {
    Mat MA, MB, R;
    Vec VI, VJ, VK;
    fill(MA, A); fill(MB, B);
    if (si) fill(VI, si);
    if (sj) fill(VJ, sj);
    if (sk) fill(VK, sk);
    int rc = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    try {
        multiply(R, C, si ? &VI : (Vec *)0, MA, tA, sj ? &VJ : (Vec *)0, MB, tB, sk ? &VK : (Vec *)0,
                 pol(policy), zero_nan != 0);
    } catch (spsparse::Exception const &) { rc = 1; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    export_mat(R, out_shape, out_n, out_i, out_k, out_v);
    return rc;
}

int ref_multiply_mv(double C, const orc_vec *si, const orc_mat *A, char tA, const orc_vec *sj,
                    const orc_vec *V, int policy, int zero_nan, uint64_t *out_shape, int64_t *out_n,
                    int32_t **out_i, double **out_v) {
    Mat MA;
    Vec VI, VJ, VV, R;
    fill(MA, A); fill(VV, V);
    if (si) fill(VI, si);
    if (sj) fill(VJ, sj);
    int rc = 0;
    try {
        multiply(R, C, si ? &VI : (Vec *)0, MA, tA, sj ? &VJ : (Vec *)0, VV, pol(policy), zero_nan != 0);
    } catch (spsparse::Exception const &) { rc = 1; }
    size_t n = R.size();
    *out_shape = R.shape[0];
    *out_n = (int64_t)n;
    *out_i = (int32_t *)malloc((n ? n : 1) * sizeof(int32_t));
    *out_v = (double *)malloc((n ? n : 1) * sizeof(double));
    for (size_t i = 0; i < n; ++i) { (*out_i)[i] = R.index(0, i); (*out_v)[i] = R.val(i); }
    return rc;
}

// Inputs of the reference's randomized multiply tests (tests/test_multiply_sparse.cpp:84-98 for
// MM, :138-152 for MV): two independent copies of std::default_random_engine(seed) drive the index
// and the value streams (std::bind copies the engine); libstdc++-specific, hence generated here
// and stored as fixtures.  Buffers need dsize*dsize slots.  mv!=0 => B is a vector (b1 unused).
void ref_testcase_inputs(unsigned dsize, int seed, int mv, int64_t *na, int32_t *a0, int32_t *a1,
                         double *av, int64_t *nb, int32_t *b0, int32_t *b1, double *bv) {
    std::default_random_engine eng(seed);
    auto draw_dim = std::bind(std::uniform_int_distribution<int>(0, dsize - 1), eng);
    auto draw_val = std::bind(std::uniform_real_distribution<double>(0, 1), eng);
    int ca = (int)(draw_val() * (double)(dsize * dsize));
    for (int t = 0; t < ca; ++t) { a0[t] = draw_dim(); a1[t] = draw_dim(); av[t] = draw_val(); }
    int cb = (int)(draw_val() * (double)(mv ? dsize : dsize * dsize));
    for (int t = 0; t < cb; ++t) { b0[t] = draw_dim(); if (!mv) b1[t] = draw_dim(); bv[t] = draw_val(); }
    *na = ca;
    *nb = cb;
}

}  // extern "C"
