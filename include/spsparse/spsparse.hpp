// <spsparse/spsparse.hpp> -- same include path as the reference; the B200 implementation lives in
// include/spsparse_b200/base.hpp.
#pragma once
#include "../spsparse_b200/base.hpp"
