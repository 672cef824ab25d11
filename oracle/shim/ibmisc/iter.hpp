// Stand-in for <ibmisc/iter.hpp>: the CRTP base used at /root/reference/slib/spsparse/array.hpp:48.
// TEST INFRASTRUCTURE ONLY (see oracle/shim/ibmisc/blitz.hpp).
#pragma once
namespace ibmisc {
template <class ValueT, class DerivedT>
struct forward_iterator {
    typedef ValueT value_type;
    bool operator!=(DerivedT const &o) const { return !(static_cast<DerivedT const &>(*this) == o); }
};
}  // namespace ibmisc
