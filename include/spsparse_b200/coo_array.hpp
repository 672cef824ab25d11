// coo_array.hpp -- the COO container and its iterators, host side.
//
// Interface mirrored (reference paths): VectorCooArray slib/spsparse/VectorCooArray.hpp:8-158 (storage
// :22-23, state :29-34, add :238-266, consolidate :299-311, dim_beginnings :323-335,
// dim_beginnings_xiter :337-344), aliases :352-356; CooIterator slib/spsparse/array.hpp:69-115;
// DimIndexIter array.hpp:47-67; DimBeginningsXiter slib/spsparse/algorithm.hpp:173-233.
// Storage stays struct-of-arrays std::vector on the host, exactly what user code indexes into; the
// heavy methods (consolidate, dim_beginnings) ship the vectors to the GPU through the C ABI.
#pragma once

#include "base.hpp"
#include "xiter.hpp"

namespace spsparse {

template <class VectorCooArrayT> class DimBeginningsXiter;
template <class IterT> class OverwriteAccum;

// ---- iterator over entries: *it is the index tuple, it.val() the value ------------------------
template <class IndicesT, class IterIndexT, int RANK, class IterValT, class CollectionT>
class CooIterator {
protected:
    CollectionT *const parent;
    int i;

public:
    static const int rank = RANK;
    typedef IndicesT indices_type;
    typedef indices_type value_type;
    typedef IterIndexT index_type;
    typedef IterValT val_type;

    CooIterator(CollectionT *p, int _i) : parent(p), i(_i) {}

    indices_type operator[](int n) { return parent->index(i + n); }
    indices_type index() { return parent->index(i); }
    indices_type operator*() { return parent->index(i); }

    CooIterator &operator+=(int n) { i += n; return *this; }
    CooIterator &operator-=(int n) { i -= n; return *this; }
    CooIterator &operator++() { ++i; return *this; }
    CooIterator &operator--() { --i; return *this; }
    CooIterator operator+(int n) const { return CooIterator(parent, i + n); }
    bool operator==(CooIterator const &rhs) const { return i == rhs.i; }
    bool operator!=(CooIterator const &rhs) const { return i != rhs.i; }

    int offset() const { return i; }
    IterIndexT &index(int k) { return parent->index(k, i); }
    void set_index(indices_type const &idx) { parent->set_index(i, idx); }
    IterValT &val() { return parent->val(i); }
};

// ---- the same walk, reporting one dimension of the index only -----------------------------------
template <class ValueT, class ValT, class IterT>
class DimIndexIter {
public:
    typedef ValueT value_type;
    IterT wrapped;
    const int dim;

    DimIndexIter(int _dim, IterT const &&ii) : wrapped(ii), dim(_dim) {}

    ValueT operator*() { return wrapped.index(dim); }
    ValT &val() { return wrapped.val(); }
    DimIndexIter &operator++() { ++wrapped; return *this; }
    bool operator==(DimIndexIter const &rhs) const { return wrapped == rhs.wrapped; }
    bool operator!=(DimIndexIter const &rhs) const { return !(wrapped == rhs.wrapped); }
};

template <class ArrayT>
std::ostream &_ostream_out_array(std::ostream &os, ArrayT const &A) {
    os << "VectorCooArray<";
    stream(os, &A.shape[0], (int)A.shape.size());
    os << ">(";
    for (auto ii(A.begin()); ii != A.end(); ++ii) {
        auto idx(ii.index());
        os << "(";
        for (int k = 0; k < A.rank; ++k) os << idx[k] << " ";
        os << ": " << ii.val() << ")";
    }
    return os << ")";
}

// =====================================================================================================
template <class IndexT, class ValT, int RANK>
class VectorCooArray {
public:
    static const int rank = RANK;
    typedef IndexT index_type;
    typedef ValT val_type;
    typedef std::array<index_type, rank> indices_type;

    std::array<size_t, RANK> shape;
    void set_shape(std::array<size_t, RANK> const &_shape) { shape = _shape; }

protected:
    typedef VectorCooArray<IndexT, ValT, RANK> ThisVectorCooArrayT;
    std::array<std::vector<IndexT>, RANK> index_vecs;
    std::vector<ValT> val_vec;
    bool dim_beginnings_set;
    std::vector<size_t> _dim_beginnings;

public:
    bool edit_mode;                    // add() is legal
    std::array<int, RANK> sort_order;  // sort_order[0] == -1: not sorted

    VectorCooArray() : dim_beginnings_set(false), edit_mode(true), sort_order() {
        sort_order[0] = -1;
        shape.fill(0);
    }
    VectorCooArray(std::array<size_t, RANK> const &_shape)
        : shape(_shape), dim_beginnings_set(false), edit_mode(true), sort_order() {
        sort_order[0] = -1;
    }
    VectorCooArray(VectorCooArray &&) = default;
    VectorCooArray(VectorCooArray const &) = default;
    void operator=(ThisVectorCooArrayT &&o) {
        shape = o.shape;
        index_vecs = std::move(o.index_vecs);
        val_vec = std::move(o.val_vec);
        dim_beginnings_set = o.dim_beginnings_set;
        _dim_beginnings = std::move(o._dim_beginnings);
        edit_mode = o.edit_mode;
        sort_order = o.sort_order;
    }
    void operator=(ThisVectorCooArrayT const &o) {
        ThisVectorCooArrayT copy(o);
        *this = std::move(copy);
    }

    std::unique_ptr<ThisVectorCooArrayT> new_blank() const { return std::unique_ptr<ThisVectorCooArrayT>(new ThisVectorCooArrayT(shape)); }
    ThisVectorCooArrayT make_blank() const { return ThisVectorCooArrayT(shape); }

    // ---- element access
    IndexT &index(int dim, size_t ix) { return index_vecs[dim][ix]; }
    IndexT const &index(int dim, size_t ix) const { return index_vecs[dim][ix]; }
    ValT &val(size_t ix) { return val_vec[ix]; }
    ValT const &val(size_t ix) const { return val_vec[ix]; }
    std::array<IndexT, RANK> index(int ix) const {
        std::array<IndexT, RANK> r;
        for (int k = 0; k < RANK; ++k) r[k] = index_vecs[k][ix];
        return r;
    }
    std::vector<IndexT> index_vec(int ix) const {
        std::vector<IndexT> r(RANK);
        for (int k = 0; k < RANK; ++k) r[k] = index_vecs[k][ix];
        return r;
    }
    void set_index(int ix, std::array<IndexT, RANK> const &idx) {
        for (int k = 0; k < RANK; ++k) index_vecs[k][ix] = idx[k];
    }

    // raw struct-of-arrays views (what the C ABI uploads from / downloads into)
    std::vector<IndexT> const &index_data(int dim) const { return index_vecs[dim]; }
    std::vector<ValT> const &val_data() const { return val_vec; }

#ifdef SPSPARSE_B200_HAVE_BLITZ
    blitz::Array<IndexT, 1> indices(int dim) const { return ibmisc::to_blitz(index_vecs[dim]); }
    blitz::Array<ValT, 1> vals() const { return ibmisc::to_blitz(val_vec); }
    void add_blitz(blitz::TinyVector<IndexT, RANK> const &index, ValT const val) {
        std::array<IndexT, RANK> ix;
        for (int k = 0; k < RANK; ++k) ix[k] = index[k];
        add(ix, val);
    }
    blitz::Array<ValT, RANK> to_dense();
#endif

    size_t size() const { return val_vec.size(); }
    void clear() {
        for (auto &v : index_vecs) v.clear();
        val_vec.clear();
        dim_beginnings_set = false;
        _dim_beginnings.clear();
        edit();
    }
    void reserve(size_t n) {
        for (auto &v : index_vecs) v.reserve(n);
        val_vec.reserve(n);
    }

    // ---- iteration
    typedef CooIterator<const std::array<IndexT, RANK>, const IndexT, RANK, const ValT, const ThisVectorCooArrayT> const_iterator;
    typedef CooIterator<std::array<IndexT, RANK>, IndexT, RANK, ValT, ThisVectorCooArrayT> iterator;
    iterator begin(int ix = 0) { return iterator(this, ix); }
    iterator end(int ix = 0) { return iterator(this, (int)size() + ix); }
    const_iterator cbegin(int ix = 0) const { return const_iterator(this, ix); }
    const_iterator cend(int ix = 0) const { return const_iterator(this, (int)size() + ix); }
    const_iterator begin(int ix = 0) const { return const_iterator(this, ix); }
    const_iterator end(int ix = 0) const { return const_iterator(this, (int)size() - ix); }

    typedef DimIndexIter<const IndexT, const ValT, const_iterator> const_dim_iterator;
    const_dim_iterator dim_iter(int dim, int ix) const { return const_dim_iterator(dim, const_iterator(this, ix)); }
    const_dim_iterator dim_begin(int dim) const { return dim_iter(dim, 0); }
    const_dim_iterator dim_end(int dim) const { return dim_iter(dim, (int)size()); }

    // ---- editing
    void edit() {
        edit_mode = true;
        sort_order[0] = -1;
    }
    void add(std::array<IndexT, RANK> const index, ValT const val) {
        if (!edit_mode) (*spsparse_error)(-1, "Must be in edit mode to use VectorCooArray::add()");
        for (int k = 0; k < RANK; ++k) {
            if (index[k] < 0 || (size_t)index[k] >= shape[k]) {
                std::ostringstream buf;
                buf << "Sparse index out of bounds: index=(";
                for (int j = 0; j < RANK; ++j) buf << index[j] << " ";
                buf << ") vs. shape=(";
                for (int j = 0; j < RANK; ++j) buf << shape[j] << " ";
                buf << ")";
                (*spsparse_error)(-1, buf.str().c_str());
            }
        }
        for (int k = 0; k < RANK; ++k) index_vecs[k].push_back(index[k]);
        val_vec.push_back(val);
    }
    void set_sorted(std::array<int, RANK> _sort_order) {
        sort_order = _sort_order;
        edit_mode = false;
    }

    // Bulk append of `n` entries straight from device results (used by consolidate/multiply when
    // every index is known to be inside `shape`); equivalent to n calls of add().
    void append_raw(size_t n, IndexT const *const *idx, ValT const *val) {
        if (!edit_mode) (*spsparse_error)(-1, "Must be in edit mode to use VectorCooArray::add()");
        for (int k = 0; k < RANK; ++k) index_vecs[k].insert(index_vecs[k].end(), idx[k], idx[k] + n);
        val_vec.insert(val_vec.end(), val, val + n);
    }

    // Room for `n` more entries at the end; returns where they start, so that a device result can be downloaded straight
    // into the array's own vectors (no intermediate copy).  The caller fills every slot.
    void grow_raw(size_t n, IndexT **idx_out, ValT **val_out) {
        if (!edit_mode) (*spsparse_error)(-1, "Must be in edit mode to use VectorCooArray::add()");
        const size_t old = val_vec.size();
        if (n >= (size_t(1) << 22)) {
            // large results: the first touch of fresh pages from ONE thread (what resize() does) runs at ~1 GB/s and was most of
            // the time of a multiply into an empty array (6 s for 14 GB); take the page faults in parallel first
            for (int k = 0; k < RANK; ++k) {
                index_vecs[k].reserve(old + n);
                spb_host_prefault(index_vecs[k].data() + old, (uint64_t)n * sizeof(IndexT));
            }
            val_vec.reserve(old + n);
            spb_host_prefault(val_vec.data() + old, (uint64_t)n * sizeof(ValT));
        }
        for (int k = 0; k < RANK; ++k) { index_vecs[k].resize(old + n); idx_out[k] = index_vecs[k].data() + old; }
        val_vec.resize(old + n);
        *val_out = val_vec.data() + old;
    }

    // ---- in-place algorithms (defined in algorithm.hpp)
    void consolidate(std::array<int, RANK> const &_sort_order, DuplicatePolicy duplicate_policy = DuplicatePolicy::ADD,
                     bool handle_nan = false);
    void transpose(std::array<int, RANK> const &sort_order);

    std::vector<size_t> const &dim_beginnings() const;
    DimBeginningsXiter<ThisVectorCooArrayT> dim_beginnings_xiter() const;
};

template <class IndexT, class ValT, int RANK>
std::ostream &operator<<(std::ostream &os, VectorCooArray<IndexT, ValT, RANK> const &A) {
    return _ostream_out_array(os, A);
}

template <class IndexT, class ValT> using VectorCooMatrix = VectorCooArray<IndexT, ValT, 2>;
template <class IndexT, class ValT> using VectorCooVector = VectorCooArray<IndexT, ValT, 1>;

// ---- row (or column) walk over a sorted array ---------------------------------------------------------
// Wraps the compressed row-start list; *it is the row's index, sub_xiter() walks the row's entries.
template <class VectorCooArrayT>
class DimBeginningsXiter : public STLXiter<std::vector<size_t>::const_iterator> {
public:
    SPSPARSE_LOCAL_TYPES(VectorCooArrayT);
    typedef std::vector<size_t>::const_iterator DimIterT;
    typedef ValSTLXiter<typename VectorCooArrayT::const_dim_iterator> sub_xiter_type;

protected:
    VectorCooArrayT const *arr;
    int index_dim, val_dim;

public:
    DimBeginningsXiter(VectorCooArrayT const *_arr, int _index_dim, int _val_dim, DimIterT const &db_begin,
                       DimIterT const &db_end)
        : STLXiter<DimIterT>(db_begin, db_end), arr(_arr), index_dim(_index_dim), val_dim(_val_dim) {}

    bool eof() { return (ii + 1) == end; }  // the last offset is the sentinel
    index_type operator*() { return arr->index(index_dim, *ii); }
    sub_xiter_type sub_xiter(int _val_dim = -1) {
        int const d = _val_dim < 0 ? val_dim : _val_dim;
        return sub_xiter_type(arr->dim_iter(d, (int)*ii), arr->dim_iter(d, (int)*(ii + 1)));
    }
};

}  // namespace spsparse
