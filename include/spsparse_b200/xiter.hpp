// xiter.hpp -- iterators that know their own end ("xiters") and sorted merge-joins over them.
//
// Interface mirrored (reference slib/spsparse/xiter.hpp): STLXiter :69-96, ValSTLXiter :110-121,
// make_val_xiter :133-136, Join3Xiter :149-194, join3_xiter :196-198, Join2Xiter :236-278,
// join2_xiter :280-282.  A join yields the values present in ALL of its ascending, non-repeating
// inputs; the caller reads the matching elements through the public sub-iterators i1, i2, i3.
// On the GPU path multiply() does not use these (the kernels merge rows directly); they are kept
// for callers that iterate on the host.
#pragma once

#include <type_traits>
#include <utility>

namespace spsparse {

template <class STLIter>
class STLXiter {
public:
    STLIter const begin;  // where iteration started
    STLIter ii;           // current position
    STLIter const end;

    typedef typename STLIter::value_type value_type;

    STLXiter(STLIter const &_begin, STLIter const &_end) : begin(_begin), ii(_begin), end(_end) {}

    bool eof() { return ii == end; }
    size_t offset() { return ii - begin; }
    void operator++() { ++ii; }
    auto operator*() -> decltype(*ii) { return *ii; }
};

// Same, for iterators that also expose val() (sparse vectors / matrix rows).
template <class ValSTLIter>
class ValSTLXiter : public STLXiter<ValSTLIter> {
public:
    ValSTLXiter(ValSTLIter const &_begin, ValSTLIter const &_end) : STLXiter<ValSTLIter>(_begin, _end) {}
    auto val() -> decltype(STLXiter<ValSTLIter>::ii.val()) { return this->ii.val(); }
};

template <class ValSTLIter>
ValSTLXiter<ValSTLIter> make_val_xiter(ValSTLIter &&_begin, ValSTLIter &&_end) {
    return ValSTLXiter<ValSTLIter>(std::move(_begin), std::move(_end));
}

namespace detail {

// Leapfrog step shared by the 2- and 3-way joins: raise `want` to the largest head seen and move
// every input up to it, until all heads agree or one input runs out.
template <class V, class X>
inline bool seek(X &x, V &want, bool &moved) {
    for (; !x.eof(); ++x) {
        V const here = *x;
        if (here == want) return true;
        if (here > want) { want = here; moved = true; return true; }
    }
    return false;  // exhausted
}

}  // namespace detail

template <class Xiter1T, class Xiter2T, class Xiter3T>
class Join3Xiter {
    typedef typename std::remove_const<typename Xiter1T::value_type>::type value_t;
    bool _eof;

    void settle() {
        if (i1.eof()) { _eof = true; return; }
        value_t want = *i1;
        for (;;) {
            bool moved = false;
            if (!detail::seek(i1, want, moved) || !detail::seek(i2, want, moved) || !detail::seek(i3, want, moved)) {
                _eof = true;
                return;
            }
            if (!moved) return;  // nobody overshot: all heads equal `want`
        }
    }

public:
    Xiter1T i1;
    Xiter2T i2;
    Xiter3T i3;
    int total_in_use;

    Join3Xiter(Xiter1T &&_i1, Xiter2T &&_i2, Xiter3T &&_i3)
        : _eof(false), i1(std::move(_i1)), i2(std::move(_i2)), i3(std::move(_i3)), total_in_use(3) {
        settle();
    }
    bool eof() { return _eof; }
    void operator++() {
        ++i1; ++i2; ++i3;
        settle();
    }
};

template <class Xiter1T, class Xiter2T, class Xiter3T>
Join3Xiter<Xiter1T, Xiter2T, Xiter3T> join3_xiter(Xiter1T &&i1, Xiter2T &&i2, Xiter3T &&i3) {
    return Join3Xiter<Xiter1T, Xiter2T, Xiter3T>(std::move(i1), std::move(i2), std::move(i3));
}

template <class Xiter1T, class Xiter2T>
class Join2Xiter {
    typedef typename std::remove_const<typename Xiter1T::value_type>::type value_t;
    bool _eof;

    void settle() {
        if (i1.eof()) { _eof = true; return; }
        value_t want = *i1;
        for (;;) {
            bool moved = false;
            if (!detail::seek(i1, want, moved) || !detail::seek(i2, want, moved)) {
                _eof = true;
                return;
            }
            if (!moved) return;
        }
    }

public:
    Xiter1T i1;
    Xiter2T i2;
    int total_in_use;

    Join2Xiter(Xiter1T &&_i1, Xiter2T &&_i2) : _eof(false), i1(std::move(_i1)), i2(std::move(_i2)), total_in_use(2) {
        settle();
    }
    bool eof() { return _eof; }
    void operator++() {
        ++i1; ++i2;
        settle();
    }
};

template <class Xiter1T, class Xiter2T>
Join2Xiter<Xiter1T, Xiter2T> join2_xiter(Xiter1T &&i1, Xiter2T &&i2) {
    return Join2Xiter<Xiter1T, Xiter2T>(std::move(i1), std::move(i2));
}

}  // namespace spsparse
