// multiply.hpp -- sparse*sparse multiply with optional diagonal scale vectors, on the GPU.
//
// Interface mirrored (reference slib/spsparse/multiply_sparse.hpp): matrix*matrix :138-248,
// matrix*vector :269-365.  Same template parameters, argument order and defaults; same observable
// steps in the same order: output shape set first, inner-dimension check through spsparse_error,
// empty short-circuit, then the product (consolidation of A/B/V included) -- which here is one call
// into the C ABI instead of the reference's row x column merge-join loops.
#pragma once

#include "algorithm.hpp"

namespace spsparse {

// ---- row walkers of one multiply operand, with or without a diagonal scale vector ------------------------
// Interface mirrored: MultXiter multiply_sparse.hpp:39-46, SimpleMultXiter :52-66, ScaledMultXiter :71-90,
// new_mult_xiter :98-111.  A walker visits the non-empty rows (leading sorted dimension) of a consolidated
// matrix; with a scale vector it visits only the rows the vector has an entry for and reports that entry.
// multiply() below does not need them -- the kernels read dense-ified scale vectors -- they are here for host
// code that walks operands the way the reference's multiply does.  The matrix must be consolidated
// (dim_beginnings_xiter needs the sort flag); the scale vector ascending and duplicate-free.
template <class MatT>
struct MultXiter {
    typedef typename DimBeginningsXiter<MatT>::sub_xiter_type sub_xiter_type;
    virtual ~MultXiter() {}
    virtual typename MatT::index_type index() = 0;       // the row's index
    virtual bool eof() = 0;
    virtual void operator++() = 0;
    virtual typename MatT::val_type scale_val() = 0;     // scale[index()], 1 without a scale vector
    virtual sub_xiter_type sub_xiter() = 0;              // the row's entries
};

template <class MatT>
class SimpleMultXiter : public MultXiter<MatT> {
    DimBeginningsXiter<MatT> rows;

public:
    explicit SimpleMultXiter(MatT const &A) : rows(A.dim_beginnings_xiter()) {}
    typename MatT::index_type index() { return *rows; }
    bool eof() { return rows.eof(); }
    void operator++() { ++rows; }
    typename MatT::val_type scale_val() { return 1; }
    typename MultXiter<MatT>::sub_xiter_type sub_xiter() { return rows.sub_xiter(); }
};

template <class MatT, class ScaleT>
class ScaledMultXiter : public MultXiter<MatT> {
    typedef ValSTLXiter<typename ScaleT::const_dim_iterator> scale_xiter_type;
    Join2Xiter<DimBeginningsXiter<MatT>, scale_xiter_type> both;  // rows present in A AND in the scale vector

public:
    ScaledMultXiter(MatT const &A, ScaleT const &scale)
        : both(A.dim_beginnings_xiter(), scale_xiter_type(scale.dim_begin(0), scale.dim_end(0))) {}
    typename MatT::index_type index() { return *both.i1; }
    bool eof() { return both.eof(); }
    void operator++() { ++both; }
    typename MatT::val_type scale_val() { return both.i2.val(); }
    typename MultXiter<MatT>::sub_xiter_type sub_xiter() { return both.i1.sub_xiter(); }
};

template <class MatT, class ScaleT>
std::unique_ptr<MultXiter<MatT>> new_mult_xiter(MatT const &A, ScaleT const *scale) {
    typedef std::unique_ptr<MultXiter<MatT>> ptr;
    return scale ? ptr(new ScaledMultXiter<MatT, ScaleT>(A, *scale)) : ptr(new SimpleMultXiter<MatT>(A));
}

// ret = C * diag(scalei) * op(A) * diag(scalej) * op(B) * diag(scalek)
template <class ScaleIT, class MatAT, class ScaleJT, class MatBT, class ScaleKT, class AccumulatorT>
void multiply(AccumulatorT &ret, double C, ScaleIT const *scalei, MatAT const &A, char transpose_A,
              ScaleJT const *scalej, MatBT const &B, char transpose_B, ScaleKT const *scalek,
              DuplicatePolicy duplicate_policy = DuplicatePolicy::ADD, bool zero_nan = false) {
    const int ar = transpose_A == 'T' ? 1 : 0, ai = 1 - ar;  // row / inner dimension of op(A)
    const int bc = transpose_B == 'T' ? 0 : 1, bi = 1 - bc;  // column / inner dimension of op(B)
    ret.set_shape({A.shape[ar], B.shape[bc]});
    if (A.shape[ai] != B.shape[bi])
        (*spsparse_error)(-1, "Inner dimensions for A (%ld) and B (%ld) must match!", (long)A.shape[ai], (long)B.shape[bi]);
    if (isnone(C) || (scalei && scalei->size() == 0) || A.size() == 0 || (scalej && scalej->size() == 0) ||
        B.size() == 0 || (scalek && scalek->size() == 0))
        return;

    b200::Handle dA, dB, dI, dJ, dK, dR;
    b200::upload(A, dA);
    b200::upload(B, dB);
    if (scalei) b200::upload(*scalei, dI);
    if (scalej) b200::upload(*scalej, dJ);
    if (scalek) b200::upload(*scalek, dK);
    b200::check(spb_multiply_mm(b200::default_context(), C, dI.h, dA.h, transpose_A, dJ.h, dB.h, transpose_B, dK.h,
                                b200::policy_code(duplicate_policy), zero_nan ? 1 : 0, &dR.h, nullptr));
    b200::deliver<2>(ret, dR.h);
}

// ret = C * diag(scalei) * op(A) * diag(scalej) * V
template <class ScaleIT, class MatAT, class ScaleJT, class VecT, class AccumulatorT>
void multiply(AccumulatorT &ret, double C, ScaleIT const *scalei, MatAT const &A, char transpose_A,
              ScaleJT const *scalej, VecT const &V, DuplicatePolicy duplicate_policy = DuplicatePolicy::ADD,
              bool zero_nan = false) {
    const int ar = transpose_A == 'T' ? 1 : 0, ai = 1 - ar;
    ret.set_shape({A.shape[ar]});
    if (A.shape[ai] != V.shape[0])
        (*spsparse_error)(-1, "Inner dimensions for A (%ld) and V (%ld) must match!", (long)A.shape[ai], (long)V.shape[0]);
    if (isnone(C) || (scalei && scalei->size() == 0) || A.size() == 0 || (scalej && scalej->size() == 0) || V.size() == 0)
        return;

    b200::Handle dA, dV, dI, dJ, dR;
    b200::upload(A, dA);
    b200::upload(V, dV);
    if (scalei) b200::upload(*scalei, dI);
    if (scalej) b200::upload(*scalej, dJ);
    b200::check(spb_multiply_mv(b200::default_context(), C, dI.h, dA.h, transpose_A, dJ.h, dV.h,
                                b200::policy_code(duplicate_policy), zero_nan ? 1 : 0, &dR.h));
    b200::deliver<1>(ret, dR.h);
}

}  // namespace spsparse
