// algorithm.hpp -- consolidate, sorted_permutation, dim_beginnings (GPU, through the C ABI) and the
// small host loops copy / transpose / to_sparse.
//
// Interface mirrored (reference slib/spsparse/algorithm.hpp): copy :30-37, transpose :46-57,
// dim_beginnings :74-118, consolidate :251-319, Consolidate<> :324-369, sorted_permutation :411-427,
// to_sparse :433-440; VectorCooArray::consolidate / dim_beginnings / dim_beginnings_xiter / to_dense
// (slib/spsparse/VectorCooArray.hpp:299-344).  First argument = output accumulator, as in the reference.
#pragma once

#include "accum.hpp"
#include "coo_array.hpp"

namespace spsparse {

namespace b200 {

// host array -> device handle (sort_order flag travels with it: VectorCooArray.hpp:33-34)
template <class ArrT>
inline void upload(ArrT const &A, Handle &out) {
    static_assert(abi_types_ok<typename ArrT::index_type, typename ArrT::val_type>::value,
                  "the B200 hot path is instantiated for 32-bit integer indices and double values "
                  "(the instantiation of every reference test)");
    const int R = ArrT::rank;
    static_assert(R == 1 || R == 2, "rank 1 and 2 arrays only");
    uint64_t shape[2] = {0, 0};
    const int32_t *idx[2] = {nullptr, nullptr};
    int so[2] = {-1, 0};
    for (int k = 0; k < R; ++k) {
        shape[k] = A.shape[k];
        idx[k] = reinterpret_cast<const int32_t *>(A.index_data(k).data());
        so[k] = A.sort_order[k];
    }
    if (so[0] < 0) so[0] = -1;
    check(spb_coo_upload(default_context(), R, shape, idx, A.val_data().data(), A.size(), so, &out.h));
}

// device result -> accumulator.  Bulk fill when the accumulator is a VectorCooArray whose shape covers
// every index; otherwise one add() per entry (same calls the reference would have made).
template <class IndexT, class ValT, int RANK>
inline bool bulk_ok(VectorCooArray<IndexT, ValT, RANK> const &ret, const uint64_t *src_shape) {
    if (!abi_types_ok<IndexT, ValT>::value) return false;   // other element types: one add() per entry
    for (int k = 0; k < RANK; ++k)
        if (ret.shape[k] < src_shape[k]) return false;
    return true;
}
template <class AccT>
inline bool bulk_ok(AccT const &, const uint64_t *) { return false; }

// the result is downloaded straight into the tail of the array's own vectors
template <class IndexT, class ValT, int RANK>
inline void bulk_append(VectorCooArray<IndexT, ValT, RANK> &ret, size_t n, spb_coo *h) {
    IndexT *ip[2] = {nullptr, nullptr};
    ValT *vp = nullptr;
    ret.grow_raw(n, ip, &vp);
    int32_t *p32[2] = {reinterpret_cast<int32_t *>(ip[0]), reinterpret_cast<int32_t *>(ip[1])};
    check(spb_coo_download(default_context(), h, p32, reinterpret_cast<double *>(vp)));
}
template <class AccT>
inline void bulk_append(AccT &, size_t, spb_coo *) {}

template <int RANK, class AccT>
inline void deliver(AccT &ret, spb_coo *h) {
    int rank = 0;
    uint64_t shape[2] = {0, 0}, n = 0;
    check(spb_coo_info(h, &rank, shape, &n, nullptr));
    if (n == 0) return;
    if (bulk_ok(ret, shape)) {
        bulk_append(ret, (size_t)n, h);
        return;
    }
    std::vector<int32_t> idx[2];
    std::vector<double> val(n);
    int32_t *ip[2] = {nullptr, nullptr};
    for (int k = 0; k < rank; ++k) { idx[k].resize(n); ip[k] = idx[k].data(); }
    check(spb_coo_download(default_context(), h, ip, val.data()));
    for (size_t t = 0; t < n; ++t) {
        std::array<int, RANK> ix;
        for (int k = 0; k < RANK; ++k) ix[k] = idx[k][t];
        ret.add(ix, val[t]);
    }
}

}  // namespace b200

// ---- copy / transpose: O(n) host loops over an accumulator --------------------------------------------
template <class VectorCooArrayT, class AccumulatorT>
void copy(AccumulatorT &ret, VectorCooArrayT const &A) {
    for (auto ii = A.begin(); ii != A.end(); ++ii) ret.add(ii.index(), ii.val());
}

// ret.dim[i] == A.dim[perm[i]]
template <class VectorCooArrayT, class AccumulatorT>
void transpose(AccumulatorT &ret, VectorCooArrayT const &A, std::array<int, VectorCooArrayT::rank> const &perm) {
    std::array<int, VectorCooArrayT::rank> idx;
    for (auto ii = A.begin(); ii != A.end(); ++ii) {
        for (int k = 0; k < VectorCooArrayT::rank; ++k) idx[k] = ii.index(perm[k]);
        ret.add(idx, ii.val());
    }
}

// ---- sorted_permutation: stable argsort by (index[sort_order[0]], index[sort_order[1]], ...) on the GPU --
template <class VectorCooArrayT>
std::vector<size_t> sorted_permutation(VectorCooArrayT const &A, std::array<int, VectorCooArrayT::rank> const &sort_order) {
    std::vector<size_t> perm(A.size());
    if (A.size() == 0) return perm;
    b200::Handle dA;
    b200::upload(A, dA);
    std::vector<uint64_t> p(A.size());
    b200::check(spb_sorted_permutation(b200::default_context(), dA.h, sort_order.data(), p.data()));
    for (size_t i = 0; i < p.size(); ++i) perm[i] = (size_t)p[i];
    return perm;
}

// ---- consolidate: sort, merge duplicates, drop zero inputs -- on the GPU --------------------------------
template <class VectorCooArrayT, class AccumulatorT>
void consolidate(AccumulatorT &ret, VectorCooArrayT const &A, std::array<int, VectorCooArrayT::rank> const &sort_order,
                 DuplicatePolicy duplicate_policy = DuplicatePolicy::ADD, bool zero_nan = false) {
    if (A.size() > 0) {
        b200::Handle dA, dR;
        b200::upload(A, dA);
        // uploaded as data to be sorted: whatever flag A carries is irrelevant to consolidate()
        b200::check(spb_consolidate(b200::default_context(), dA.h, sort_order.data(), b200::policy_code(duplicate_policy),
                                    zero_nan ? 1 : 0, &dR.h, nullptr));
        b200::deliver<VectorCooArrayT::rank>(ret, dR.h);
    }
    ret.set_sorted(sort_order);
}

// Consolidates only if A is not already flagged sorted in the requested order; never touches A.
template <class ArrayT>
class Consolidate {
    static const int rank = ArrayT::rank;
    std::unique_ptr<ArrayT> A2;
    ArrayT const *Ap;

public:
    Consolidate(ArrayT const *A, std::array<int, ArrayT::rank> const &sort_order,
                DuplicatePolicy duplicate_policy = DuplicatePolicy::ADD, bool zero_nan = false) {
        if (A->sort_order == sort_order) {
            Ap = A;
        } else {
            A2 = A->new_blank();
            Ap = A2.get();
            consolidate(*A2, *A, sort_order, duplicate_policy, zero_nan);
        }
    }
    ArrayT const &operator()() { return *Ap; }
};

// ---- dim_beginnings: offsets where the leading sorted index changes, plus the sentinel -- on the GPU ----
template <class VectorCooArrayT>
std::vector<size_t> dim_beginnings(VectorCooArrayT const &A) {
    std::vector<size_t> out;
    if (A.sort_order[0] < 0) (*spsparse_error)(-1, "dim_beginnings() required the VectorCooArray is sorted first.");
    if (A.size() == 0) return out;
    b200::Handle dA;
    b200::upload(A, dA);
    std::vector<uint64_t> buf(A.size() + 1);
    uint64_t count = 0;
    b200::check(spb_dim_beginnings(b200::default_context(), dA.h, buf.data(), buf.size(), &count));
    out.assign(buf.begin(), buf.begin() + count);
    return out;
}

// ---- methods of VectorCooArray that need the algorithms -----------------------------------------------------
template <class IndexT, class ValT, int RANK>
void VectorCooArray<IndexT, ValT, RANK>::consolidate(std::array<int, RANK> const &_sort_order,
                                                     DuplicatePolicy duplicate_policy, bool handle_nan) {
    if (this->sort_order == _sort_order && !edit_mode) return;  // already consolidated this way
    ThisVectorCooArrayT ret(shape);
    spsparse::consolidate(ret, *this, _sort_order, duplicate_policy, handle_nan);
    *this = std::move(ret);
}

template <class IndexT, class ValT, int RANK>
void VectorCooArray<IndexT, ValT, RANK>::transpose(std::array<int, RANK> const &perm) {
    OverwriteAccum<iterator> overwrite(begin());
    spsparse::transpose(overwrite, *this, perm);
}

template <class IndexT, class ValT, int RANK>
std::vector<size_t> const &VectorCooArray<IndexT, ValT, RANK>::dim_beginnings() const {
    if (!dim_beginnings_set) {  // lazy, cached until clear() / assignment
        ThisVectorCooArrayT *self = const_cast<ThisVectorCooArrayT *>(this);
        self->_dim_beginnings = spsparse::dim_beginnings(*this);
        self->dim_beginnings_set = true;
    }
    return _dim_beginnings;
}

template <class IndexT, class ValT, int RANK>
DimBeginningsXiter<VectorCooArray<IndexT, ValT, RANK>> VectorCooArray<IndexT, ValT, RANK>::dim_beginnings_xiter() const {
    auto &db(dim_beginnings());
    return DimBeginningsXiter<ThisVectorCooArrayT>(this, sort_order[0], sort_order[1], db.begin(), db.end());
}

#ifdef SPSPARSE_B200_HAVE_BLITZ
template <class IndexT, class ValT, int RANK>
blitz::Array<ValT, RANK> VectorCooArray<IndexT, ValT, RANK>::to_dense() {
    blitz::TinyVector<int, RANK> ext;
    for (int k = 0; k < RANK; ++k) ext[k] = (int)shape[k];
    blitz::Array<ValT, RANK> ret(ext);
    ret = 0;
    DenseAccum<ThisVectorCooArrayT> accum(ret);
    copy(accum, *this);
    return ret;
}

template <class TypeT, int RANK, class CooArrayT>
void to_sparse(CooArrayT &ret, blitz::Array<TypeT, RANK> const &arr) {
    for (auto ii = arr.begin(); ii != arr.end(); ++ii)
        if (*ii != 0) ret.add_blitz(ii.position(), *ii);
}
#endif

}  // namespace spsparse
