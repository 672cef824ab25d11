// accum.hpp -- output accumulators: anything with add(index, val).
//
// Interface mirrored (reference slib/spsparse/accum.hpp): OverwriteAccum :43-57, PermuteAccum :73-93,
// DenseAccum :110-140 (needs blitz), ScalarAccumulator :158-167.  consolidate() and multiply()
// deliver GPU results to an arbitrary accumulator through add(); when the accumulator is a
// VectorCooArray they fill its vectors in bulk instead.
#pragma once

#include "base.hpp"

namespace spsparse {

// Writes over the entries of an existing array, in order (e.g. in-place transpose).
template <class IterT>
class OverwriteAccum {
    SPSPARSE_LOCAL_TYPES(IterT);
    IterT ii;

public:
    OverwriteAccum(IterT &&_ii) : ii(std::move(_ii)) {}
    void add(indices_type const &index, typename IterT::val_type const &val) {
        ii.set_index(index);
        ii.val() = val;
        ++ii;
    }
};

// Reorders / selects dimensions on the way into another accumulator.
template <int IN_RANK, class AccumulatorT>
class PermuteAccum {
public:
    static const int rank = IN_RANK;
    static const int out_rank = AccumulatorT::rank;

private:
    AccumulatorT sub;
    std::vector<int> perm;
    std::array<int, out_rank> out_idx;

public:
    PermuteAccum(AccumulatorT &&_sub, std::vector<int> const &_perm) : sub(std::move(_sub)), perm(_perm) {}
    void add(std::array<int, IN_RANK> const &index, typename AccumulatorT::val_type const &val) {
        for (int k = 0; k < out_rank; ++k) out_idx[k] = index[perm[k]];
        sub.add(out_idx, val);
    }
};

#ifdef SPSPARSE_B200_HAVE_BLITZ
// Accumulates into a dense blitz array (copies of blitz arrays share storage).
template <class VectorCooArrayT>
struct DenseAccum {
    SPSPARSE_LOCAL_TYPES(VectorCooArrayT);
    typedef blitz::Array<val_type, rank> blitz_type;

private:
    DuplicatePolicy duplicate_policy;
    blitz_type dense;
    blitz::TinyVector<int, rank> bidx;

public:
    DenseAccum(blitz_type &_dense, DuplicatePolicy _duplicate_policy = DuplicatePolicy::ADD)
        : duplicate_policy(_duplicate_policy), dense(_dense) {}
    void add(indices_type const &index, val_type const &val) {
        for (int k = 0; k < rank; ++k) bidx[k] = (int)index[k];
        val_type &cell(dense(bidx));
        if (duplicate_policy == DuplicatePolicy::ADD) cell += val;
        else if (duplicate_policy == DuplicatePolicy::REPLACE) cell = val;
        else if (!std::isnan(cell)) cell = val;  // LEAVE_ALONE, as written in accum.hpp:128-130
    }
};
#endif

// Sums every value, ignoring the indices.
template <class VectorCooArrayT>
struct ScalarAccumulator {
    SPSPARSE_LOCAL_TYPES(VectorCooArrayT);
    val_type val;
    ScalarAccumulator() : val(0) {}
    void add(const std::array<index_type, rank> &, val_type const &_val) { val += _val; }
};

}  // namespace spsparse
