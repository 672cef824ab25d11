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
