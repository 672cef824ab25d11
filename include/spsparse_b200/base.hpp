// base.hpp -- enums, error hook, sort-order constants and the `isnone` predicate of namespace
// spsparse, plus the process-wide GPU context the template layer runs on.
//
// Interface mirrored (reference paths): DuplicatePolicy slib/spsparse/spsparse.hpp:25-26,
// Exception :30-38, error_ptr / spsparse_error :47,54 (default handler spsparse.cpp:12-28),
// SPSPARSE_LOCAL_TYPES :73-78, ROW_MAJOR / COL_MAJOR :82-83 (spsparse.cpp:30-31), isnone :95-103.
// The objects (spsparse_error, ROW_MAJOR, COL_MAJOR) live in libspsparse_b200.so, as they live in
// libspsparse.so for the reference.
#pragma once

#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <exception>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../spsparse_b200.h"

#if defined(__has_include)
#if __has_include(<ibmisc/blitz.hpp>)
#include <ibmisc/blitz.hpp>
#define SPSPARSE_B200_HAVE_BLITZ 1
#endif
#endif

namespace spsparse {

enum class DuplicatePolicy { LEAVE_ALONE, ADD, REPLACE };

class Exception : public std::exception {
public:
    virtual ~Exception() {}
    virtual const char *what() const noexcept { return "spsparse::Exception()"; }
};

// printf-style error hook; the default prints to stderr and throws spsparse::Exception.  Callers do
// not expect it to return.
typedef void (*error_ptr)(int retcode, char const *str, ...);
extern error_ptr spsparse_error;

extern const std::array<int, 2> ROW_MAJOR;
extern const std::array<int, 2> COL_MAJOR;

#define SPSPARSE_LOCAL_TYPES(ArrayOrIterT)                     \
    static const int rank = ArrayOrIterT::rank;                \
    typedef typename ArrayOrIterT::index_type index_type;      \
    typedef typename ArrayOrIterT::val_type val_type;          \
    typedef std::array<index_type, rank> indices_type

template <class NumT>
inline bool isnone(NumT const n, bool const zero_nan = false) {
    return (n == 0) || (zero_nan && std::isnan(n));
}

namespace b200 {

// The C ABI speaks int32 indices and double values; these are the only instantiations the reference's
// own tests use (tests/test_multiply_sparse.cpp:90-97).
template <class IndexT, class ValT>
struct abi_types_ok {
    static const bool value = std::is_integral<IndexT>::value && sizeof(IndexT) == 4 && std::is_same<ValT, double>::value;
};

inline int policy_code(DuplicatePolicy p) {
    return p == DuplicatePolicy::LEAVE_ALONE ? SPB_LEAVE_ALONE : (p == DuplicatePolicy::ADD ? SPB_ADD : SPB_REPLACE);
}

// Routes a failed C-ABI call into the library's error convention (no CPU fallback: a CUDA failure
// is an error like any other).
inline void check(int rc) {
    if (rc != SPB_OK) (*spsparse_error)(-1, "%s", spb_last_error());
}

// One context (device + stream + memory pool) per process, created on first use.  The device comes
// from the environment variable SPSPARSE_B200_DEVICE (default 0).
spb_ctx *default_context();

// RAII for device handles inside the templates
struct Handle {
    spb_coo *h = nullptr;
    Handle() {}
    Handle(Handle const &) = delete;
    Handle &operator=(Handle const &) = delete;
    ~Handle() { if (h) spb_coo_free(default_context(), h); }
};

}  // namespace b200
}  // namespace spsparse

// Writes {a, b, c} for an index tuple (used by operator<< of the arrays).
template <class T>
std::ostream &stream(std::ostream &os, T const *const a, int RANK) {
    os << "{";
    for (int k = 0; k < RANK; ++k) os << (k ? ", " : "") << a[k];
    return os << "}";
}
