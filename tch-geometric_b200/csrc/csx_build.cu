// edge_index (COO) -> CSC / CSR on sm_100a, and ind2ptr.
//
// Replaces src/data/storage.rs:67-101 (ind2ptr) and :103-127 (argsort(col*N + row), ind2ptr,
// gathers) of the reference.  The sort key packs (major << minor_bits) | minor, which orders
// exactly like major*size_minor + minor because minor < size_minor <= 2^minor_bits, and only the
// key's significant bits are radix-sorted (44 bits for the products-shaped graph instead of 64).
// The value carried through the sort is the 32-bit edge id (perm); `indices` and the major ids are
// decoded from the sorted keys, so no random row[perm] / col[perm] gathers are needed.
// The device-wide radix sort itself is CUB's DeviceRadixSort (library code, like cuBLAS for a plain
// GEMM); key build, decode and ind2ptr are the kernels below.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace tchgeo {
namespace {

constexpr int CSX_THREADS = 256;

__host__ __device__ inline int bits_for(int64_t n) {  // bits needed for values in [0, n)
  int b = 0;
  while (b < 63 && ((int64_t)1 << b) < n) ++b;
  return b < 1 ? 1 : b;
}

__global__ void __launch_bounds__(CSX_THREADS) build_keys_kernel(const int64_t* __restrict__ major,
                                                                const int64_t* __restrict__ minor, int64_t E,
                                                                int64_t n_major, int64_t n_minor, int minor_bits,
                                                                uint64_t* __restrict__ keys,
                                                                uint32_t* __restrict__ vals, uint32_t* err) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    const int64_t a = major[e], b = minor[e];
    if (a < 0 || a >= n_major || b < 0 || b >= n_minor) atomicOr(err, DEV_ERR_INDEX);
    keys[e] = ((uint64_t)a << minor_bits) | (uint64_t)b;
    vals[e] = (uint32_t)e;
  }
}

// sorted keys -> indices (minor ids), perm (i64) and ptrs (ind2ptr of the major ids, storage.rs:67-101)
__global__ void __launch_bounds__(CSX_THREADS) decode_kernel(const uint64_t* __restrict__ keys,
                                                            const uint32_t* __restrict__ vals, int64_t E,
                                                            int64_t n_major, int minor_bits,
                                                            int64_t* __restrict__ ptrs, int64_t* __restrict__ indices,
                                                            int64_t* __restrict__ perm) {
  const uint64_t mask = (minor_bits >= 64) ? ~0ull : ((1ull << minor_bits) - 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    const uint64_t key = keys[e];
    const int64_t maj = (int64_t)(key >> minor_bits);
    indices[e] = (int64_t)(key & mask);
    perm[e] = (int64_t)vals[e];
    // out[i] = e for every i in (major[e-1], major[e]]; out[0..=major[0]] = 0
    int64_t lo = (e == 0) ? 0 : (int64_t)(keys[e - 1] >> minor_bits) + 1;
    if (maj >= n_major) continue;  // flagged by build_keys_kernel
    for (int64_t i = lo; i <= maj; ++i) ptrs[i] = e;
    if (e == E - 1)
      for (int64_t i = maj + 1; i <= n_major; ++i) ptrs[i] = E;
  }
}

__global__ void __launch_bounds__(CSX_THREADS) ind2ptr_kernel(const int64_t* __restrict__ ind, int64_t numel,
                                                             int64_t m, int64_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (numel == 0) {  // storage.rs:78-80
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= m; i += stride) out[i] = 0;
    return;
  }
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel; e += stride) {
    const int64_t cur = ind[e];
    int64_t lo = (e == 0) ? 0 : ind[e - 1] + 1;
    if (lo < 0) lo = 0;
    const int64_t hi = cur < m ? cur : m;
    for (int64_t i = lo; i <= hi; ++i) out[i] = e;
    if (e == numel - 1)
      for (int64_t i = (cur < -1 ? -1 : cur) + 1; i <= m; ++i) out[i] = numel;
  }
}

__global__ void fill_zero_i64_kernel(int64_t* p, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = 0;
}

struct CsxLayout {
  size_t off_err, off_keys0, off_keys1, off_vals0, off_vals1, off_cub, cub_bytes, total;
};

inline size_t align_up(size_t x) { return (x + 255) / 256 * 256; }

cudaError_t csx_layout(int64_t E, int end_bit, CsxLayout& L) {
  size_t cub_bytes = 0;
  cub::DoubleBuffer<uint64_t> dk(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, dk, dv, (int64_t)(E > 0 ? E : 1), 0, end_bit);
  if (e != cudaSuccess) return e;
  const size_t n = (size_t)(E > 0 ? E : 1);
  L.off_err = 0;
  L.off_keys0 = 256;
  L.off_keys1 = L.off_keys0 + align_up(n * 8);
  L.off_vals0 = L.off_keys1 + align_up(n * 8);
  L.off_vals1 = L.off_vals0 + align_up(n * 4);
  L.off_cub = L.off_vals1 + align_up(n * 4);
  L.cub_bytes = cub_bytes;
  L.total = L.off_cub + align_up(cub_bytes) + 256;
  return cudaSuccess;
}

inline unsigned grid_for(int64_t n) {
  int64_t g = (n + CSX_THREADS - 1) / CSX_THREADS;
  const int64_t cap = 148 * 16;  // grid-stride loops: 16 resident 256-thread CTAs' worth per SM
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace
}  // namespace tchgeo

using namespace tchgeo;

extern "C" tchgeo_status tchgeo_ind2ptr(const int64_t* ind, int64_t numel, int64_t m, int64_t* out,
                                        tchgeo_stream stream_) {
  TCHGEO_REQUIRE(numel >= 0 && m >= 0 && out != nullptr && (numel == 0 || ind != nullptr), "bad ind2ptr argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  ind2ptr_kernel<<<grid_for(numel > 0 ? numel : m + 1), CSX_THREADS, 0, stream>>>(ind, numel, m, out);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  return TCHGEO_OK;
}

extern "C" size_t tchgeo_coo_to_csx_workspace_bytes(int64_t num_edges, int64_t n_rows, int64_t n_cols) {
  if (num_edges < 0 || n_rows < 0 || n_cols < 0) return 0;
  CsxLayout L;
  if (csx_layout(num_edges, 64, L) != cudaSuccess) return 0;
  return L.total;
}

extern "C" tchgeo_status tchgeo_coo_to_csx(const int64_t* row, const int64_t* col, int64_t E, int64_t n_rows,
                                           int64_t n_cols, int32_t csc, int64_t* ptrs, int64_t* indices, int64_t* perm,
                                           void* workspace, size_t workspace_bytes, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(E >= 0 && n_rows >= 0 && n_cols >= 0, "negative size");
  TCHGEO_REQUIRE(E < ((int64_t)1 << 32), "more than 2^32-1 edges are not supported");
  TCHGEO_REQUIRE(ptrs != nullptr && (E == 0 || (row && col && indices && perm)), "NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t n_major = csc ? n_cols : n_rows;
  const int64_t n_minor = csc ? n_rows : n_cols;
  const int64_t* major = csc ? col : row;
  const int64_t* minor = csc ? row : col;
  if (E == 0) {  // storage.rs:78-80
    fill_zero_i64_kernel<<<grid_for(n_major + 1), CSX_THREADS, 0, stream>>>(ptrs, n_major + 1);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
    return TCHGEO_OK;
  }
  const int minor_bits = bits_for(n_minor), major_bits = bits_for(n_major);
  TCHGEO_REQUIRE(minor_bits + major_bits <= 64, "graph too large: (row, col) key needs more than 64 bits");
  const int end_bit = minor_bits + major_bits;
  CsxLayout L;
  TCHGEO_CUDA_CHECK(csx_layout(E, end_bit, L));
  TCHGEO_REQUIRE(workspace != nullptr, "workspace is NULL");
  if (workspace_bytes < L.total) {
    set_last_error("workspace too small: need %zu bytes, got %zu", L.total, workspace_bytes);
    return TCHGEO_ERR_CAPACITY;
  }
  char* ws = (char*)workspace;
  uint32_t* err = (uint32_t*)(ws + L.off_err);
  uint64_t* keys0 = (uint64_t*)(ws + L.off_keys0);
  uint64_t* keys1 = (uint64_t*)(ws + L.off_keys1);
  uint32_t* vals0 = (uint32_t*)(ws + L.off_vals0);
  uint32_t* vals1 = (uint32_t*)(ws + L.off_vals1);
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(err, 0, 256, stream));
  build_keys_kernel<<<grid_for(E), CSX_THREADS, 0, stream>>>(major, minor, E, n_major, n_minor, minor_bits, keys0,
                                                             vals0, err);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  cub::DoubleBuffer<uint64_t> dk(keys0, keys1);
  cub::DoubleBuffer<uint32_t> dv(vals0, vals1);
  size_t cub_bytes = L.cub_bytes;
  TCHGEO_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ws + L.off_cub, cub_bytes, dk, dv, E, 0, end_bit, stream));
  decode_kernel<<<grid_for(E), CSX_THREADS, 0, stream>>>(dk.Current(), dv.Current(), E, n_major, minor_bits, ptrs,
                                                         indices, perm);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  uint32_t herr = 0;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  return status_from_dev_err(herr);
}
