// Microbenchmark: DRAM bytes per random 8-byte gather for different load flavours on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_modes gather_modes.cu
// Run under: ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum ./gather_modes
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

enum { LD_PLAIN, LD_NC, LD_NC_NA, LD_CG, LD_CS, LD_CV, LD_LU, LD_NC_EF, LD_NC_L2_64, LD_I32_NC, N_MODES };
const char* names[] = {"ld.global", "ld.global.nc", "ld.global.nc.L1::no_allocate", "ld.global.cg", "ld.global.cs",
                       "ld.global.cv", "ld.global.lu", "ld.global.nc + L2 evict_first policy", "ld.global.nc.L2::64B",
                       "ld.global.nc.s32 (4-byte table)"};

template <int MODE>
__device__ __forceinline__ int64_t load(const int64_t* p, uint64_t pol) {
  int64_t v;
  if (MODE == LD_PLAIN) asm volatile("ld.global.s64 %0, [%1];" : "=l"(v) : "l"(p));
  if (MODE == LD_NC) asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(p));
  if (MODE == LD_NC_NA) asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(p));
  if (MODE == LD_CG) asm volatile("ld.global.cg.s64 %0, [%1];" : "=l"(v) : "l"(p));
  if (MODE == LD_CS) asm volatile("ld.global.cs.s64 %0, [%1];" : "=l"(v) : "l"(p));
  if (MODE == LD_CV) asm volatile("ld.global.cv.s64 %0, [%1];" : "=l"(v) : "l"(p));
  if (MODE == LD_LU) asm volatile("ld.global.lu.s64 %0, [%1];" : "=l"(v) : "l"(p));
  if (MODE == LD_NC_EF) asm volatile("ld.global.nc.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
  if (MODE == LD_NC_L2_64) asm volatile("ld.global.nc.L2::64B.s64 %0, [%1];" : "=l"(v) : "l"(p));
  if (MODE == LD_I32_NC) {
    int32_t w;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(w) : "l"(reinterpret_cast<const int32_t*>(p)));
    v = w;
  }
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(256) gather(const int64_t* __restrict__ table, const uint32_t* __restrict__ idx,
                                              int64_t* __restrict__ out, int64_t n) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t* p = MODE == LD_I32_NC ? reinterpret_cast<const int64_t*>(reinterpret_cast<const int32_t*>(table) + idx[i])
                                       : table + idx[i];
  const int64_t v = load<MODE>(p, pol);
  asm volatile("st.global.cs.s64 [%0], %1;" ::"l"(out + i), "l"(v) : "memory");
}

__global__ void fill(int64_t* t, int64_t n, uint32_t* idx, int64_t m, uint32_t nmod) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) t[i] = i;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    uint64_t x = (uint64_t)i * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    idx[i] = (uint32_t)(x % nmod);
  }
}

template <int MODE>
void run(const int64_t* t, const uint32_t* idx, int64_t* out, int64_t m) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  gather<MODE><<<(unsigned)((m + 255) / 256), 256>>>(t, idx, out, m);
  cudaEventRecord(a);
  gather<MODE><<<(unsigned)((m + 255) / 256), 256>>>(t, idx, out, m);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("%-42s %8.3f ms  %6.2f G gathers/s\n", names[MODE], ms, m / ms / 1e6);
}

int main() {
  const int64_t n = 61859140, m = 64 * 1000 * 1000;
  int64_t *t, *out; uint32_t* idx;
  cudaMalloc(&t, n * 8); cudaMalloc(&out, m * 8); cudaMalloc(&idx, m * 4);
  fill<<<148 * 8, 256>>>(t, n, idx, m, (uint32_t)n);
  cudaDeviceSynchronize();
  run<LD_PLAIN>(t, idx, out, m); run<LD_NC>(t, idx, out, m); run<LD_NC_NA>(t, idx, out, m); run<LD_CG>(t, idx, out, m);
  run<LD_CS>(t, idx, out, m); run<LD_CV>(t, idx, out, m); run<LD_LU>(t, idx, out, m); run<LD_NC_EF>(t, idx, out, m);
  run<LD_NC_L2_64>(t, idx, out, m); run<LD_I32_NC>(t, idx, out, m);
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
  printf("-- after cudaLimitMaxL2FetchGranularity = 32\n");
  run<LD_PLAIN>(t, idx, out, m); run<LD_NC_NA>(t, idx, out, m);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
