// <spsparse/netcdf.hpp> -- same include path as the reference; the implementation lives in
// include/spsparse_b200/netcdf.hpp.  Needs <netcdf> and <ibmisc/netcdf.hpp>: the real libraries, or the
// minimal stand-ins of include/spsparse_b200/mini_netcdf/ (add that directory to the include path).
#pragma once
#include "../spsparse_b200/netcdf.hpp"
