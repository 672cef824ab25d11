// host_symbols.cpp -- the objects of namespace spsparse that the reference keeps in its one compiled
// file (slib/spsparse/spsparse.cpp:12-31): the default error handler, the replaceable error hook and the
// ROW_MAJOR / COL_MAJOR constants; plus the process-wide GPU context of the template layer.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "../../include/spsparse_b200/base.hpp"

namespace spsparse {

// prints like printf, then throws: callers rely on it not returning (spsparse.cpp:12-26)
static void default_error(int /*retcode*/, const char *format, ...) {
    va_list ap;
    va_start(ap, format);
    vfprintf(stderr, format, ap);
    va_end(ap);
    fprintf(stderr, "\n");
    throw spsparse::Exception();
}

error_ptr spsparse_error = &default_error;
const std::array<int, 2> ROW_MAJOR = {0, 1};
const std::array<int, 2> COL_MAJOR = {1, 0};

namespace b200 {

spb_ctx *default_context() {
    static spb_ctx *ctx = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *d = getenv("SPSPARSE_B200_DEVICE");
        if (spb_ctx_create(d ? atoi(d) : 0, nullptr, &ctx) != SPB_OK) ctx = nullptr;
    });
    if (!ctx) (*spsparse_error)(-1, "%s", spb_last_error());
    return ctx;
}

}  // namespace b200
}  // namespace spsparse
