// abi.cu — error plumbing and version query of the C-ABI (include/dccf_b200.h).
#include <stdarg.h>
#include "common.cuh"

namespace dccf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
}  // namespace dccf

extern "C" const char* dccf_last_error(void) { return dccf::g_err; }
extern "C" int dccf_abi_version(void) { return DCCF_ABI_VERSION; }
