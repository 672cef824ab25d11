/*
 * spsparse_oracle.c -- CPU restatement of the spsparse hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity oracle for the B200 build.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path (spsparse_b200/, include/) never links or calls it.
 *
 * It restates, in plain C, the algorithms of the reference (paths relative to /root/reference):
 *   - sorted_permutation / CmpIndex          slib/spsparse/algorithm.hpp:375-427
 *   - consolidate                             slib/spsparse/algorithm.hpp:251-319
 *   - dim_beginnings                          slib/spsparse/algorithm.hpp:74-118
 *   - Join2Xiter / Join3Xiter merge-join      slib/spsparse/xiter.hpp:149-278, next_noincr_body.hpp:1-53
 *   - multiply (matrix*matrix)                slib/spsparse/multiply_sparse.hpp:152-248
 *   - multiply (matrix*vector)                slib/spsparse/multiply_sparse.hpp:281-365
 *   - isnone                                  slib/spsparse/spsparse.hpp:95-103
 *   - transpose / copy                        slib/spsparse/algorithm.hpp:30-57
 *   - to_dense (DenseAccum) / to_sparse       slib/spsparse/VectorCooArray.hpp:313-321, accum.hpp:110-140,
 *                                             algorithm.hpp:433-440
 *
 * Parity pinning: tests/test_oracle_*.py check every function here against (a) the golden
 * vectors of the reference's own tests (tests/test_array.cpp:67-79,135-168,
 * tests/test_xiter.cpp:52-125, tests/test_multiply_sparse.cpp:41-79) and (b) outputs of the
 * genuine reference headers compiled into oracle/_ref (fixtures in tests/golden/).
 *
 * Two multiply formulations are given on purpose:
 *   orc_multiply_mm_pairs : the reference's own loop nest (row x column pairs, merge-join) --
 *                           quadratic, only for small cases; it is the literal restatement.
 *   orc_multiply_mm       : row-wise (Gustavson) evaluation that obeys the same rules and adds
 *                           the terms of every output in the same ascending-j order, hence gives
 *                           bit-identical results; usable at the sizes the reference cannot reach.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 * Types: IndexT=int32, ValT=double, RANK in {1,2}  (SURVEY.md section 8).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_LEAVE_ALONE = 0, ORC_ADD = 1, ORC_REPLACE = 2 }; /* spsparse.hpp:25-26 (same order) */

/* spsparse.hpp:95-103 */
static int isnone(double v, int zero_nan) {
    if (zero_nan) return isnan(v) || (v == 0);
    return (v == 0);
}

/* ------------------------------------------------------------------------------------------
 * sorted_permutation: stable argsort by (idx[so[0]], idx[so[1]], ...)   algorithm.hpp:375-427
 * std::stable_sort is any stable comparison sort; a bottom-up merge sort is used here.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int rank;
    const int32_t *idx[2];
    int so[2];
} cmp_ctx;

/* CmpIndex::operator()  algorithm.hpp:386-395 */
static int cmp_less(const cmp_ctx *c, int64_t i, int64_t j) {
    for (int k = 0; k < c->rank - 1; ++k) {
        const int32_t *d = c->idx[c->so[k]];
        if (d[i] < d[j]) return 1;
        if (d[i] > d[j]) return 0;
    }
    const int32_t *d = c->idx[c->so[c->rank - 1]];
    return d[i] < d[j];
}

static void stable_argsort(const cmp_ctx *c, int64_t n, int64_t *perm) {
    int64_t *tmp = (int64_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    int64_t *src = perm, *dst = tmp;
    for (int64_t i = 0; i < n; ++i) perm[i] = i; /* algorithm.hpp:419-421 */
    for (int64_t w = 1; w < n; w *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * w) {
            int64_t mid = lo + w < n ? lo + w : n;
            int64_t hi = lo + 2 * w < n ? lo + 2 * w : n;
            int64_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi) {
                /* take from the right run only if strictly less: keeps equal keys in input order */
                if (cmp_less(c, src[b], src[a])) dst[o++] = src[b++];
                else dst[o++] = src[a++];
            }
            while (a < mid) dst[o++] = src[a++];
            while (b < hi) dst[o++] = src[b++];
        }
        int64_t *t = src; src = dst; dst = t;
    }
    if (src != perm) memcpy(perm, src, (size_t)n * sizeof(int64_t));
    free(tmp);
}

void orc_sorted_permutation(int rank, int64_t n, const int32_t *idx0, const int32_t *idx1,
                            const int *sort_order, int64_t *perm) {
    cmp_ctx c;
    c.rank = rank;
    c.idx[0] = idx0;
    c.idx[1] = idx1;
    c.so[0] = sort_order[0];
    c.so[1] = rank > 1 ? sort_order[1] : 0;
    stable_argsort(&c, n, perm);
}

/* ------------------------------------------------------------------------------------------
 * consolidate   algorithm.hpp:251-319
 * Outputs must have room for n entries.  Returns the number of entries written.
 * ---------------------------------------------------------------------------------------- */
int64_t orc_consolidate(int rank, int64_t n, const int32_t *idx0, const int32_t *idx1,
                        const double *val, const int *sort_order, int policy, int zero_nan,
                        int32_t *out0, int32_t *out1, double *outv) {
    if (n <= 0) return 0; /* :263 */
    int64_t *perm = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    orc_sorted_permutation(rank, n, idx0, idx1, sort_order, perm); /* :266 */
    const int32_t *idx[2] = {idx0, idx1};
    int32_t *out[2] = {out0, out1};
    int64_t m = 0, ii = 0;

    /* leading run: skip 0 (and NaN when zero_nan)  :272-275 */
    for (;; ++ii) {
        if (ii == n) goto finished;
        if (!isnone(val[perm[ii]], zero_nan)) break;
    }
    {
        int32_t acc_idx[2] = {0, 0};
        for (int k = 0; k < rank; ++k) acc_idx[k] = idx[k][perm[ii]]; /* :278 own dim order */
        double acc_val = val[perm[ii]];                               /* :279 */
        ++ii;
        for (;; ++ii) {
            /* later entries: zeros dropped, NaN kept (zero_nan NOT passed)  :284-292 */
            for (;; ++ii) {
                if (ii == n) {
                    for (int k = 0; k < rank; ++k) out[k][m] = acc_idx[k]; /* :287 */
                    outv[m] = acc_val;
                    ++m;
                    goto finished;
                }
                if (!isnone(val[perm[ii]], 0)) break;
            }
            int same = 1;
            for (int k = 0; k < rank; ++k)
                if (idx[k][perm[ii]] != acc_idx[k]) { same = 0; break; } /* :295-304 */
            if (!same) {
                for (int k = 0; k < rank; ++k) out[k][m] = acc_idx[k]; /* :299 written unconditionally */
                outv[m] = acc_val;
                ++m;
                for (int k = 0; k < rank; ++k) acc_idx[k] = idx[k][perm[ii]];
                acc_val = val[perm[ii]];
            } else if (policy == ORC_ADD) {
                acc_val += val[perm[ii]]; /* :307-308 left fold, insertion order */
            } else if (policy == ORC_REPLACE) {
                acc_val = val[perm[ii]]; /* :309-310 */
            } /* LEAVE_ALONE keeps the first */
        }
    }
finished:
    free(perm);
    return m;
}

/* ------------------------------------------------------------------------------------------
 * transpose   algorithm.hpp:46-57: every entry, in order, with its indices permuted
 * (new dimension k takes old dimension perm[k]); copy (:30-37) is the identity permutation.
 * ---------------------------------------------------------------------------------------- */
void orc_transpose(int rank, int64_t n, const int32_t *idx0, const int32_t *idx1, const int *perm,
                   int32_t *out0, int32_t *out1) {
    const int32_t *idx[2] = {idx0, idx1};
    int32_t *out[2] = {out0, out1};
    for (int64_t i = 0; i < n; ++i)
        for (int new_k = 0; new_k < rank; ++new_k) out[new_k][i] = idx[perm[new_k]][i]; /* :51-54 */
}

/* ------------------------------------------------------------------------------------------
 * to_dense   VectorCooArray.hpp:313-321: a zeroed row-major array, then copy() (algorithm.hpp:30-37)
 * into a DenseAccum (accum.hpp:110-140), entry after entry in stored order.  The method itself uses ADD;
 * the accumulator's other two policies are restated as written (LEAVE_ALONE overwrites unless the cell
 * holds a NaN, :128-130).  Returns -1 if an index is outside the shape (blitz would not check).
 * ---------------------------------------------------------------------------------------- */
int orc_to_dense(int rank, const uint64_t *shape, int64_t n, const int32_t *idx0, const int32_t *idx1,
                 const double *val, int policy, double *dense) {
    uint64_t cells = 1;
    for (int k = 0; k < rank; ++k) cells *= shape[k];
    for (uint64_t c = 0; c < cells; ++c) dense[c] = 0.0; /* :316 ret = 0 */
    for (int64_t i = 0; i < n; ++i) {
        if (idx0[i] < 0 || (uint64_t)idx0[i] >= shape[0]) return -1;
        uint64_t c = (uint64_t)idx0[i];
        if (rank == 2) {
            if (idx1[i] < 0 || (uint64_t)idx1[i] >= shape[1]) return -1;
            c = c * shape[1] + (uint64_t)idx1[i];
        }
        double *oval = &dense[c];
        if (policy == ORC_LEAVE_ALONE) { if (!isnan(*oval)) *oval = val[i]; } /* accum.hpp:128-130 */
        else if (policy == ORC_ADD) *oval += val[i];                           /* :131-133 */
        else *oval = val[i];                                                   /* :134-136 */
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * to_sparse   algorithm.hpp:433-440: every element != 0 (NaN included) of a row-major dense array,
 * in storage order.  Outputs need room for every cell.  Returns the number of entries.
 * ---------------------------------------------------------------------------------------- */
int64_t orc_to_sparse(int rank, const uint64_t *shape, const double *dense, int32_t *out0, int32_t *out1,
                      double *outv) {
    uint64_t cells = 1;
    for (int k = 0; k < rank; ++k) cells *= shape[k];
    int64_t m = 0;
    for (uint64_t c = 0; c < cells; ++c) {
        if (dense[c] != 0) { /* :438 */
            if (rank == 2) { out0[m] = (int32_t)(c / shape[1]); out1[m] = (int32_t)(c % shape[1]); }
            else out0[m] = (int32_t)c;
            outv[m] = dense[c];
            ++m;
        }
    }
    return m;
}

/* ------------------------------------------------------------------------------------------
 * dim_beginnings   algorithm.hpp:74-118.  idx_dim = index vector of dimension sort_order[0].
 * out needs room for n+1.  Returns the number of offsets written (0 for an empty array).
 * ---------------------------------------------------------------------------------------- */
int64_t orc_dim_beginnings(int64_t n, const int32_t *idx_dim, int64_t *out) {
    int64_t m = 0;
    if (n <= 0) return 0; /* :89 */
    out[m++] = 0;         /* :90 */
    int32_t last = idx_dim[0];
    for (int64_t i = 1;; ++i) {
        if (i == n) { out[m++] = n; break; } /* sentinel :95-98 */
        if (idx_dim[i] != last) { out[m++] = i; last = idx_dim[i]; }
    }
    return m;
}

/* ------------------------------------------------------------------------------------------
 * Merge-join of ascending, non-repeating lists   next_noincr_body.hpp:1-53, xiter.hpp:164-192,251-276
 * A tiny state machine that mirrors the reference exactly (including its `next_match` logic),
 * so that misuse cases (unsorted / repeating input) behave like the reference too.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const int32_t *v[3];
    int64_t pos[3], end[3];
    int nlist;
    int32_t next_match;
    int eof;
} joiner;

static void join_next_noincr(joiner *J) {
restart:
    for (;; ++J->pos[0]) { /* next_noincr_body.hpp:5-15 */
        if (J->pos[0] == J->end[0]) { J->eof = 1; return; }
        int32_t x = J->v[0][J->pos[0]];
        if (x == J->next_match) break;
        if (x > J->next_match) { J->next_match = x; break; }
    }
    if (J->nlist >= 2) {
        for (;; ++J->pos[1]) { /* :20-31 */
            if (J->pos[1] == J->end[1]) { J->eof = 1; return; }
            int32_t x = J->v[1][J->pos[1]];
            if (x == J->next_match) break;
            if (x > J->next_match) { J->next_match = x; ++J->pos[0]; goto restart; }
        }
    }
    if (J->nlist >= 3) {
        for (;; ++J->pos[2]) { /* :36-48 */
            if (J->pos[2] == J->end[2]) { J->eof = 1; return; }
            int32_t x = J->v[2][J->pos[2]];
            if (x == J->next_match) break;
            if (x > J->next_match) { J->next_match = x; ++J->pos[0]; ++J->pos[1]; goto restart; }
        }
    }
}

static void join_init(joiner *J, int nlist, const int32_t *a, int64_t a0, int64_t a1,
                      const int32_t *b, int64_t b0, int64_t b1, const int32_t *c, int64_t c0,
                      int64_t c1) {
    J->nlist = nlist;
    J->v[0] = a; J->pos[0] = a0; J->end[0] = a1;
    J->v[1] = b; J->pos[1] = b0; J->end[1] = b1;
    J->v[2] = c; J->pos[2] = c0; J->end[2] = c1;
    J->next_match = 0;
    J->eof = (a0 == a1); /* xiter.hpp:170,255 */
    if (J->eof) return;
    J->next_match = a[a0];
    join_next_noincr(J);
}

static void join_incr(joiner *J) { /* xiter.hpp:185-192, 270-276 */
    for (int k = 0; k < J->nlist; ++k) ++J->pos[k];
    join_next_noincr(J);
}

/* Lists the matching values of a 2- or 3-way join (nc<0 => 2-way). Returns the count. */
int64_t orc_join(const int32_t *a, int64_t na, const int32_t *b, int64_t nb, const int32_t *c,
                 int64_t nc, int32_t *out) {
    joiner J;
    int64_t m = 0;
    join_init(&J, nc < 0 ? 2 : 3, a, 0, na, b, 0, nb, c, 0, nc < 0 ? 0 : nc);
    for (; !J.eof; join_incr(&J)) out[m++] = a[J.pos[0]];
    return m;
}

/* ------------------------------------------------------------------------------------------
 * A small owning COO container for internal use.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t n;
    int32_t *idx[2];
    double *val;
    int owned;
} coo;

static void coo_free(coo *c) {
    if (c->owned) { free(c->idx[0]); free(c->idx[1]); free(c->val); }
    c->n = 0;
}

/* Consolidate<>   algorithm.hpp:354-369: reuse the input iff its sort_order flag equals the
 * requested one (edit_mode is not looked at), else consolidate into a temporary. */
static coo consolidate_if_needed(int rank, int64_t n, const int32_t *idx0, const int32_t *idx1,
                                 const double *val, const int *cur_so, const int *want_so,
                                 int policy, int zero_nan) {
    coo r;
    int same = 1;
    for (int k = 0; k < rank; ++k) if (cur_so[k] != want_so[k]) same = 0;
    if (same) {
        r.n = n; r.idx[0] = (int32_t *)idx0; r.idx[1] = (int32_t *)idx1; r.val = (double *)val;
        r.owned = 0;
        return r;
    }
    size_t cap = (size_t)(n > 0 ? n : 1);
    r.idx[0] = (int32_t *)malloc(cap * sizeof(int32_t));
    r.idx[1] = (int32_t *)malloc(cap * sizeof(int32_t));
    r.val = (double *)malloc(cap * sizeof(double));
    r.owned = 1;
    r.n = orc_consolidate(rank, n, idx0, idx1, val, want_so, policy, zero_nan, r.idx[0], r.idx[1],
                          r.val);
    return r;
}

/* growable output */
typedef struct {
    int64_t n, cap;
    int32_t *i0, *i1;
    double *v;
} outbuf;

static void out_push(outbuf *o, int32_t a, int32_t b, double v) {
    if (o->n == o->cap) {
        o->cap = o->cap ? o->cap * 2 : 1024;
        o->i0 = (int32_t *)realloc(o->i0, (size_t)o->cap * sizeof(int32_t));
        o->i1 = (int32_t *)realloc(o->i1, (size_t)o->cap * sizeof(int32_t));
        o->v = (double *)realloc(o->v, (size_t)o->cap * sizeof(double));
    }
    o->i0[o->n] = a; o->i1[o->n] = b; o->v[o->n] = v; ++o->n;
}

void orc_free(void *p) { free(p); }

/* Matrix operand as the caller holds it: shape, entries, and the array's sort_order flag
 * ({-1,x} when unsorted / in edit mode, VectorCooArray.hpp:121-125). */
typedef struct {
    uint64_t shape[2];
    int64_t n;
    const int32_t *idx0, *idx1;
    const double *val;
    int sort_order[2];
} orc_mat;

/* Sparse vector operand (scale vectors / V).  Scale vectors are used as stored
 * (multiply_sparse.hpp:83-86,225): ascending, non-repeating. */
typedef struct {
    uint64_t shape;
    int64_t n;
    const int32_t *idx;
    const double *val;
    int sort_order; /* -1 unsorted, 0 sorted */
} orc_vec;

enum { ORC_OK = 0, ORC_ERR_INNER_DIM = 1 };

/* ------------------------------------------------------------------------------------------
 * multiply, matrix*matrix -- literal loop nest   multiply_sparse.hpp:152-248
 * ---------------------------------------------------------------------------------------- */
int orc_multiply_mm_pairs(double C, const orc_vec *si, const orc_mat *A, char tA, const orc_vec *sj,
                          const orc_mat *B, char tB, const orc_vec *sk, int policy, int zero_nan,
                          uint64_t out_shape[2], int64_t *out_n, int32_t **out_i, int32_t **out_k,
                          double **out_v) {
    static const int ROW_MAJOR[2] = {0, 1}, COL_MAJOR[2] = {1, 0};
    const int *a_so = (tA == 'T') ? COL_MAJOR : ROW_MAJOR; /* :167 */
    const int *b_so = (tB == 'T') ? ROW_MAJOR : COL_MAJOR; /* :168 */
    out_shape[0] = A->shape[a_so[0]];                       /* :169 */
    out_shape[1] = B->shape[b_so[0]];
    *out_n = 0; *out_i = NULL; *out_k = NULL; *out_v = NULL;
    if (A->shape[a_so[1]] != B->shape[b_so[1]]) return ORC_ERR_INNER_DIM; /* :172-174 */
    if (isnone(C, 0) || (si && si->n == 0) || A->n == 0 || (sj && sj->n == 0) || B->n == 0 ||
        (sk && sk->n == 0))
        return ORC_OK; /* :178-184 */

    coo Ac = consolidate_if_needed(2, A->n, A->idx0, A->idx1, A->val, A->sort_order, a_so, policy,
                                   zero_nan); /* :187 */
    coo Bc = consolidate_if_needed(2, B->n, B->idx0, B->idx1, B->val, B->sort_order, b_so, policy,
                                   zero_nan); /* :188 */
    outbuf o = {0, 0, NULL, NULL, NULL};
    if (Ac.n == 0 || Bc.n == 0) goto done; /* (the reference would misbehave) */
    {
        int64_t *ab = (int64_t *)malloc((size_t)(Ac.n + 1) * sizeof(int64_t));
        int64_t *bb = (int64_t *)malloc((size_t)(Bc.n + 1) * sizeof(int64_t));
        int64_t na = orc_dim_beginnings(Ac.n, Ac.idx[a_so[0]], ab) - 1; /* rows, sentinel excluded */
        int64_t nb = orc_dim_beginnings(Bc.n, Bc.idx[b_so[0]], bb) - 1;
        const int32_t *arow = Ac.idx[a_so[0]], *aj = Ac.idx[a_so[1]];
        const int32_t *bcol = Bc.idx[b_so[0]], *bj = Bc.idx[b_so[1]];
        /* row heads / col heads as lists, so ScaledMultXiter's Join2 (:79-92) can be restated */
        int32_t *rows = (int32_t *)malloc((size_t)(na > 0 ? na : 1) * sizeof(int32_t));
        int32_t *cols = (int32_t *)malloc((size_t)(nb > 0 ? nb : 1) * sizeof(int32_t));
        for (int64_t r = 0; r < na; ++r) rows[r] = arow[ab[r]];
        for (int64_t c = 0; c < nb; ++c) cols[c] = bcol[bb[c]];

        joiner JA;
        int64_t ra = 0;
        if (si) join_init(&JA, 2, rows, 0, na, si->idx, 0, si->n, NULL, 0, 0);
        for (;;) { /* Loop 1  :192-193 */
            int64_t r;
            double a_scale;
            if (si) { if (JA.eof) break; r = JA.pos[0]; a_scale = si->val[JA.pos[1]]; }
            else { if (ra == na) break; r = ra; a_scale = 1; }
            if (!isnone(a_scale, 0)) { /* :195 */
                joiner JB;
                int64_t cb = 0;
                if (sk) join_init(&JB, 2, cols, 0, nb, sk->idx, 0, sk->n, NULL, 0, 0);
                for (;;) { /* Loop 2  :208-209 */
                    int64_t c;
                    double b_scale;
                    if (sk) { if (JB.eof) break; c = JB.pos[0]; b_scale = sk->val[JB.pos[1]]; }
                    else { if (cb == nb) break; c = cb; b_scale = 1; }
                    if (!isnone(b_scale, 0)) { /* :211 */
                        double sum = 0;        /* :219 */
                        joiner J;
                        if (sj) { /* :223-228 */
                            join_init(&J, 3, aj, ab[r], ab[r + 1], sj->idx, 0, sj->n, bj, bb[c],
                                      bb[c + 1]);
                            for (; !J.eof; join_incr(&J))
                                sum += Ac.val[J.pos[0]] * sj->val[J.pos[1]] * Bc.val[J.pos[2]];
                        } else { /* :231-235 */
                            join_init(&J, 2, aj, ab[r], ab[r + 1], bj, bb[c], bb[c + 1], NULL, 0, 0);
                            for (; !J.eof; join_incr(&J)) sum += Ac.val[J.pos[0]] * Bc.val[J.pos[1]];
                        }
                        if (!isnone(sum, 0)) /* :238-242 */
                            out_push(&o, rows[r], cols[c], sum * C * a_scale * b_scale);
                    }
                    if (sk) join_incr(&JB); else ++cb;
                }
            }
            if (si) join_incr(&JA); else ++ra;
        }
        free(ab); free(bb); free(rows); free(cols);
    }
done:
    coo_free(&Ac); coo_free(&Bc);
    *out_n = o.n; *out_i = o.i0; *out_k = o.i1; *out_v = o.v;
    return ORC_OK;
}

/* membership lookup in an ascending, non-repeating list: returns position or -1 */
static int64_t find_sorted(const int32_t *v, int64_t n, int32_t x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) / 2;
        if (v[mid] < x) lo = mid + 1; else hi = mid;
    }
    return (lo < n && v[lo] == x) ? lo : -1;
}

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* ------------------------------------------------------------------------------------------
 * multiply, matrix*matrix -- row-wise evaluation of the same contract.
 * Every output (i,k) receives its terms in ascending j, starting from sum=0, each term formed
 * as (a*s)*b or a*b exactly as multiply_sparse.hpp:228,235 -- hence bit-identical sums.
 * Optional stats: stats[0]=F (intermediate products), stats[1]=nnz(Acon), stats[2]=nnz(Bcon),
 * stats[3]=rows of op(A) that are non-empty.
 * ---------------------------------------------------------------------------------------- */
int orc_multiply_mm(double C, const orc_vec *si, const orc_mat *A, char tA, const orc_vec *sj,
                    const orc_mat *B, char tB, const orc_vec *sk, int policy, int zero_nan,
                    uint64_t out_shape[2], int64_t *out_n, int32_t **out_i, int32_t **out_k,
                    double **out_v, int64_t *stats) {
    static const int ROW_MAJOR[2] = {0, 1}, COL_MAJOR[2] = {1, 0};
    const int *a_so = (tA == 'T') ? COL_MAJOR : ROW_MAJOR; /* M1 */
    const int *b_so = (tB == 'T') ? ROW_MAJOR : COL_MAJOR;
    out_shape[0] = A->shape[a_so[0]]; /* M2 */
    out_shape[1] = B->shape[b_so[0]];
    *out_n = 0; *out_i = NULL; *out_k = NULL; *out_v = NULL;
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (A->shape[a_so[1]] != B->shape[b_so[1]]) return ORC_ERR_INNER_DIM; /* M3 */
    if (isnone(C, 0) || (si && si->n == 0) || A->n == 0 || (sj && sj->n == 0) || B->n == 0 ||
        (sk && sk->n == 0))
        return ORC_OK; /* M4 */

    coo Ac = consolidate_if_needed(2, A->n, A->idx0, A->idx1, A->val, A->sort_order, a_so, policy,
                                   zero_nan); /* M5 */
    coo Bc = consolidate_if_needed(2, B->n, B->idx0, B->idx1, B->val, B->sort_order, b_so, policy,
                                   zero_nan);
    outbuf o = {0, 0, NULL, NULL, NULL};
    if (Ac.n == 0 || Bc.n == 0) goto done;
    {
        const uint64_t nj = A->shape[a_so[1]], nk = B->shape[b_so[0]];
        const int32_t *arow = Ac.idx[a_so[0]], *aj = Ac.idx[a_so[1]];
        const int32_t *bcol = Bc.idx[b_so[0]], *bj = Bc.idx[b_so[1]];
        /* Re-bucket Bcon by inner index j (stable => ascending k inside each j). */
        int64_t *bptr = (int64_t *)calloc((size_t)nj + 1, sizeof(int64_t));
        for (int64_t e = 0; e < Bc.n; ++e) bptr[(uint64_t)bj[e] + 1]++;
        for (uint64_t j = 0; j < nj; ++j) bptr[j + 1] += bptr[j];
        int32_t *bk = (int32_t *)malloc((size_t)Bc.n * sizeof(int32_t));
        double *bv = (double *)malloc((size_t)Bc.n * sizeof(double));
        {
            int64_t *fill = (int64_t *)malloc((size_t)(nj ? nj : 1) * sizeof(int64_t));
            memcpy(fill, bptr, (size_t)nj * sizeof(int64_t));
            for (int64_t e = 0; e < Bc.n; ++e) {
                int64_t p = fill[bj[e]]++;
                bk[p] = bcol[e];
                bv[p] = Bc.val[e];
            }
            free(fill);
        }
        /* per-column scale (M7): dense copy; absent or 0 => column excluded */
        double *skd = NULL;
        if (sk) {
            skd = (double *)calloc((size_t)(nk ? nk : 1), sizeof(double));
            for (int64_t t = 0; t < sk->n; ++t) /* entries beyond the result's width can never join */
                if (sk->idx[t] >= 0 && (uint64_t)sk->idx[t] < nk) skd[sk->idx[t]] = sk->val[t];
        }
        double *acc = (double *)malloc((size_t)(nk ? nk : 1) * sizeof(double));
        int64_t *stamp = (int64_t *)malloc((size_t)(nk ? nk : 1) * sizeof(int64_t));
        for (uint64_t k = 0; k < nk; ++k) stamp[k] = -1;
        int32_t *touched = (int32_t *)malloc((size_t)(nk ? nk : 1) * sizeof(int32_t));
        int64_t F = 0, nrows = 0;

        for (int64_t e0 = 0; e0 < Ac.n;) { /* M6: ascending non-empty rows of op(A) */
            int64_t e1 = e0 + 1;
            while (e1 < Ac.n && arow[e1] == arow[e0]) ++e1;
            int32_t i = arow[e0];
            ++nrows;
            double a_scale = 1;
            int use = 1;
            if (si) {
                int64_t p = find_sorted(si->idx, si->n, i);
                if (p < 0) use = 0; else { a_scale = si->val[p]; if (isnone(a_scale, 0)) use = 0; }
            }
            if (use) {
                int64_t nt = 0;
                for (int64_t e = e0; e < e1; ++e) { /* ascending j  (M8) */
                    int32_t j = aj[e];
                    double as = Ac.val[e];
                    if (sj) {
                        int64_t p = find_sorted(sj->idx, sj->n, j);
                        if (p < 0) continue;
                        as = as * sj->val[p]; /* (a*s) first  :228 */
                    }
                    for (int64_t q = bptr[j]; q < bptr[(uint64_t)j + 1]; ++q) {
                        int32_t k = bk[q];
                        ++F;
                        if (stamp[k] != e0) { stamp[k] = e0; acc[k] = 0; touched[nt++] = k; }
                        acc[k] += as * bv[q];
                    }
                }
                qsort(touched, (size_t)nt, sizeof(int32_t), cmp_i32); /* M10: k ascending */
                for (int64_t t = 0; t < nt; ++t) {
                    int32_t k = touched[t];
                    double b_scale = 1;
                    if (sk) { b_scale = skd[k]; if (isnone(b_scale, 0)) continue; }
                    if (!isnone(acc[k], 0)) out_push(&o, i, k, acc[k] * C * a_scale * b_scale); /* M9 */
                }
            }
            e0 = e1;
        }
        if (stats) { stats[0] = F; stats[1] = Ac.n; stats[2] = Bc.n; stats[3] = nrows; }
        free(bptr); free(bk); free(bv); free(skd); free(acc); free(stamp); free(touched);
    }
done:
    coo_free(&Ac); coo_free(&Bc);
    *out_n = o.n; *out_i = o.i0; *out_k = o.i1; *out_v = o.v;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * multiply, matrix*vector   multiply_sparse.hpp:281-365  (literal: one merge-join per row)
 * ---------------------------------------------------------------------------------------- */
int orc_multiply_mv(double C, const orc_vec *si, const orc_mat *A, char tA, const orc_vec *sj,
                    const orc_vec *V, int policy, int zero_nan, uint64_t *out_shape, int64_t *out_n,
                    int32_t **out_i, double **out_v) {
    static const int ROW_MAJOR[2] = {0, 1}, COL_MAJOR[2] = {1, 0};
    const int *a_so = (tA == 'T') ? COL_MAJOR : ROW_MAJOR; /* :294 */
    *out_shape = A->shape[a_so[0]];                         /* :295 */
    *out_n = 0; *out_i = NULL; *out_v = NULL;
    if (A->shape[a_so[1]] != V->shape) return ORC_ERR_INNER_DIM; /* :298-300 */
    if (isnone(C, 0) || (si && si->n == 0) || A->n == 0 || (sj && sj->n == 0) || V->n == 0)
        return ORC_OK; /* :304-309 */
    coo Ac = consolidate_if_needed(2, A->n, A->idx0, A->idx1, A->val, A->sort_order, a_so, policy,
                                   zero_nan); /* :312 */
    const int v_want[2] = {0, 0}, v_cur[2] = {V->sort_order, 0};
    coo Vc = consolidate_if_needed(1, V->n, V->idx, NULL, V->val, v_cur, v_want, policy, zero_nan); /* :313 */
    outbuf o = {0, 0, NULL, NULL, NULL};
    if (Ac.n == 0) goto done;
    {
        int64_t *ab = (int64_t *)malloc((size_t)(Ac.n + 1) * sizeof(int64_t));
        int64_t na = orc_dim_beginnings(Ac.n, Ac.idx[a_so[0]], ab) - 1;
        const int32_t *arow = Ac.idx[a_so[0]], *aj = Ac.idx[a_so[1]];
        int32_t *rows = (int32_t *)malloc((size_t)(na > 0 ? na : 1) * sizeof(int32_t));
        for (int64_t r = 0; r < na; ++r) rows[r] = arow[ab[r]];
        joiner JA;
        int64_t ra = 0;
        if (si) join_init(&JA, 2, rows, 0, na, si->idx, 0, si->n, NULL, 0, 0);
        for (;;) { /* :319-320 */
            int64_t r;
            double a_scale;
            if (si) { if (JA.eof) break; r = JA.pos[0]; a_scale = si->val[JA.pos[1]]; }
            else { if (ra == na) break; r = ra; a_scale = 1; }
            if (!isnone(a_scale, 0)) { /* :322 */
                double sum = 0;         /* :334 */
                joiner J;
                if (sj) { /* :338-343 */
                    join_init(&J, 3, aj, ab[r], ab[r + 1], sj->idx, 0, sj->n, Vc.idx[0], 0, Vc.n);
                    for (; !J.eof; join_incr(&J))
                        sum += Ac.val[J.pos[0]] * sj->val[J.pos[1]] * Vc.val[J.pos[2]];
                } else { /* :346-353 */
                    join_init(&J, 2, aj, ab[r], ab[r + 1], Vc.idx[0], 0, Vc.n, NULL, 0, 0);
                    for (; !J.eof; join_incr(&J)) sum += Ac.val[J.pos[0]] * Vc.val[J.pos[1]];
                }
                if (!isnone(sum, 0)) out_push(&o, rows[r], 0, sum * C * a_scale); /* :356-360 */
            }
            if (si) join_incr(&JA); else ++ra;
        }
        free(ab); free(rows);
    }
done:
    coo_free(&Ac); coo_free(&Vc);
    *out_n = o.n; *out_i = o.i0; *out_v = o.v;
    free(o.i1);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic-input generator shared by tests and bench: splitmix64
 * finaliser on a counter; identical arithmetic to the device generators in spsparse_b200/csrc.
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t x) { return (double)(mix64(x) >> 11) * 0x1.0p-53; }

/* config 2: N entries over `ubase` distinct draws, shape 2^bits x 2^bits, ~30% duplicates.
 * zero_every>0 additionally zeroes values where mix((S^0x2E80)+i)%zero_every==0. */
void orc_gen_dup_coo(uint64_t seed, int64_t i0, int64_t n, int64_t ubase, int bits,
                     int64_t zero_every, int32_t *row, int32_t *col, double *val) {
    const uint64_t mask = (1ull << bits) - 1;
    for (int64_t t = 0; t < n; ++t) {
        uint64_t i = (uint64_t)(i0 + t);
        uint64_t s = (int64_t)i < ubase ? i : mix64((seed ^ 0xD0B1Eull) + i) % (uint64_t)ubase;
        row[t] = (int32_t)(mix64(seed + 2 * s) & mask);
        col[t] = (int32_t)(mix64(seed + 2 * s + 1) & mask);
        double v = 0.5 + u01((seed ^ 0xA11CEull) + i);
        if (zero_every > 0 && mix64((seed ^ 0x2E80ull) + i) % (uint64_t)zero_every == 0) v = 0.0;
        val[t] = v;
    }
}
