#include "common.cuh"
#include <cstdarg>
#include "../../include/spaa_b200.h"

namespace spaa {
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace spaa

extern "C" {
const char* spaa_last_error(void) { return spaa::g_err; }
int spaa_abi_version(void) { return 1; }
}
