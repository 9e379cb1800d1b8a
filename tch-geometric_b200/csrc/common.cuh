// Shared device/host helpers for the sm_100a sampling library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tchgeo_cuda.h"

namespace tchgeo {

// ---- error plumbing ---------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define TCHGEO_CUDA_CHECK(expr)                                                                   \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::tchgeo::set_last_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,               \
                               cudaGetErrorString(_e));                                           \
      return TCHGEO_ERR_CUDA;                                                                     \
    }                                                                                             \
  } while (0)

#define TCHGEO_REQUIRE(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::tchgeo::set_last_error(__VA_ARGS__);      \
      return TCHGEO_ERR_BAD_ARG;                  \
    }                                             \
  } while (0)

// device-side error codes OR-ed into the workspace error word
enum : uint32_t {
  DEV_ERR_CAPACITY = 1u,
  DEV_ERR_INDEX = 2u,
  DEV_ERR_PANIC = 4u,
  DEV_ERR_WATCHDOG = 8u,
};

inline tchgeo_status status_from_dev_err(uint32_t e) {
  if (e == 0) return TCHGEO_OK;
  if (e & DEV_ERR_WATCHDOG) {
    set_last_error("internal: look-back watchdog tripped");
    return TCHGEO_ERR_INTERNAL;
  }
  if (e & DEV_ERR_INDEX) {
    set_last_error("node id out of range (the reference panics on this input)");
    return TCHGEO_ERR_INDEX;
  }
  if (e & DEV_ERR_PANIC) {
    set_last_error("input on which the reference panics (fanout 0 with non-empty neighbourhood, or non-positive weight sum)");
    return TCHGEO_ERR_REFERENCE_PANIC;
  }
  set_last_error("output capacity exceeded");
  return TCHGEO_ERR_CAPACITY;
}

// RNG stream tags (DESIGN.md "RNG contract"; mirrored by the oracle's counter mode)
constexpr uint32_t TAG_RESERVOIR = 1u;
constexpr uint32_t TAG_REPLACE = 2u;
constexpr uint32_t TAG_WEIGHTED = 3u;
constexpr uint32_t TAG_WALK = 4u;

#ifdef __CUDACC__
// ---- Philox4x32-10 (Salmon et al. SC'11), counter-based, stateless ----------------------------
struct Philox4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ uint32_t pick4(const Philox4& r, uint32_t i) {
  return i == 0 ? r.x : (i == 1 ? r.y : (i == 2 ? r.z : r.w));
}

// ---- relaxed gpu-scope 64-bit accesses for the decoupled look-back status words ---------------
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// L2 cache policies (createpolicy): the colptr array is small and re-read all the time -> evict_last;
// random single-use gathers must not push it out -> evict_first.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ int64_t ld_nc_l2hint_i64(const int64_t* p, uint64_t pol) {
  int64_t v;
  asm volatile("ld.global.nc.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int64_t ld_nc_na_l2hint_i64(const int64_t* p, uint64_t pol) {
  int64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int32_t ld_nc_na_l2hint_i32(const int32_t* p, uint64_t pol) {
  int32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

// Random gathers with a 64-byte L2 fetch.  Measured on B200 (tools/micro/gather_modes.cu): a plain
// ld.global / ld.global.nc miss fetches ~111 B from DRAM per random 8-byte gather (the L2 promotes the miss
// to the whole 128-byte line), L1::no_allocate / .cs / evict_first variants fetch 124 B and are 13 % slower,
// while the .L2::64B qualifier brings it down to 62 B.  cudaLimitMaxL2FetchGranularity has no effect.
__device__ __forceinline__ int64_t ld_gather64_i64(const int64_t* p) {
  int64_t v;
  asm volatile("ld.global.nc.L2::64B.s64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int32_t ld_gather64_i32(const int32_t* p) {
  int32_t v;
  asm volatile("ld.global.nc.L2::64B.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int64_t ld_gather64_keep_i64(const int64_t* p, uint64_t pol) {
  int64_t v;
  asm volatile("ld.global.nc.L2::cache_hint.L2::64B.s64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
  return v;
}

// streaming (evict-first) 8-byte store for write-once outputs
__device__ __forceinline__ void st_cs_i64(int64_t* p, int64_t v) {
  asm volatile("st.global.cs.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// read-only 8-byte gather that does not allocate in L1 (random single-use sector)
__device__ __forceinline__ int64_t ld_nc_na_i64(const int64_t* p) {
  int64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
#endif  // __CUDACC__

}  // namespace tchgeo
