// api.cu -- error reporting and version of the C ABI (include/snb200.h).
#include <stdarg.h>
#include "common.cuh"

namespace snb {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace snb

extern "C" const char *snb_last_error(void) { return snb::g_err; }
extern "C" int32_t snb_version(void) { return 100; }
