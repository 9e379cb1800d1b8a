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

// Hint for the current device's L2 -> DRAM fetch granularity (cudaLimitMaxL2FetchGranularity: 32, 64 or
// 128 bytes).  The sampling path is dominated by random 8-byte gathers, each of which needs one 32-byte
// sector; larger fetch granularities read neighbouring sectors that are never used.
extern "C" tchgeo_status tchgeo_device_set_l2_fetch_granularity(int32_t bytes, int32_t* actual) {
  TCHGEO_REQUIRE(bytes == 32 || bytes == 64 || bytes == 128, "granularity must be 32, 64 or 128");
  TCHGEO_CUDA_CHECK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
  size_t v = 0;
  TCHGEO_CUDA_CHECK(cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity));
  if (actual) *actual = (int32_t)v;
  return TCHGEO_OK;
}

extern "C" tchgeo_status tchgeo_status_from_error_word(uint32_t word) { return tchgeo::status_from_dev_err(word); }
