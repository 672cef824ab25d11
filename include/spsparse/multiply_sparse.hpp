// <spsparse/multiply_sparse.hpp> -- same include path as the reference; the B200 implementation lives in
// include/spsparse_b200/multiply.hpp.
#pragma once
#include "../spsparse_b200/multiply.hpp"
