// ABI bookkeeping: version and the thread-local last-error string.
#include <stdarg.h>

#include "common.cuh"

namespace tchgeo {
namespace {
thread_local char g_last_error[512] = "";
}
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
}  // namespace tchgeo

extern "C" int32_t tchgeo_abi_version(void) { return TCHGEO_ABI_VERSION; }
extern "C" const char* tchgeo_last_error(void) { return tchgeo::g_last_error; }
