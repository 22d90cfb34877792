// Error string, version.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace effimvs {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("EFFIMVS_PDL");
        return e ? atoi(e) != 0 : false;
    }();
    return on;
}
}  // namespace effimvs

extern "C" const char* effimvs_last_error(void) { return effimvs::g_err; }
extern "C" int effimvs_version(void) { return 100; }  // 0.1.0
