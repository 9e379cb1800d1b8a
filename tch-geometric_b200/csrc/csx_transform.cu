// Per-column transforms of CSC edge data on sm_100a.
//
// Replaces src/data/transform.rs:36-60 (csc_edge_cumsum) and :7-34 (csc_sort_edges) of the reference.
//
// csc_edge_cumsum: in-place inclusive prefix sum of `row_data` inside every column.  The sums are taken
// SERIALLY in CSC order (acc = acc + x, transform.rs:53-57) because floating-point addition is not
// associative: a parallel scan would round differently.  Serial order is also exactly the w_sum sequence of
// reservoir_sampling_weighted (src/utils/sampling.rs:37-48), which is what lets the weighted hop kernel read
// w_sum_i from this array instead of scanning the weights (neighbor_sampling.cu, `weights_cumsum`).
// One thread per column; a warp's 32 columns are adjacent in memory, so the lines a warp touches are
// reused from L1 across iterations.  One-time precompute per weight tensor: HBM-bound, 16 B per edge.
//
// csc_sort_edges: new_perm[col_start + i] = perm[col_start + argsort(weights[col])[i]]: a segmented
// stable sort of (weight, perm) pairs by weight.  The sort is CUB's DeviceSegmentedSort (library code, as the
// radix sort of to_csc); the reference's torch argsort is unstable, so ties are unspecified there and
// resolved here by CSC position.
#include <cub/device/device_segmented_sort.cuh>

#include "common.cuh"

namespace tchgeo {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) cumsum_kernel(const int64_t* __restrict__ col_ptrs, int64_t n_cols,
                                                     T* __restrict__ row_data, int64_t numel, uint32_t* err) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  const int64_t s = col_ptrs[c];
  int64_t e = col_ptrs[c + 1];
  if (e - s <= 1) return;          // transform.rs:46-48
  if (s < 0) { atomicOr(err, DEV_ERR_INDEX); return; }
  if (e > numel) e = numel;        // Tensor::slice clamps the end (the reference's own KAT relies on it)
  T acc = T(0);
  for (int64_t p = s; p < e; ++p) {
    acc = acc + row_data[p];
    row_data[p] = acc;
  }
}

// Tensor::slice clamps both ends into [0, numel] (the reference's own KATs end with a pointer past numel)
__global__ void __launch_bounds__(256) clamp_ptrs_kernel(const int64_t* __restrict__ src, int64_t n, int64_t numel,
                                                         int64_t* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = min(max(src[i], (int64_t)0), numel);
}

inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace
}  // namespace tchgeo

using namespace tchgeo;

extern "C" tchgeo_status tchgeo_csc_edge_cumsum_f64(const int64_t* col_ptrs, int64_t n_cols, double* row_data,
                                                    int64_t numel, int32_t* scratch, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(n_cols >= 0 && numel >= 0, "negative size");
  TCHGEO_REQUIRE(col_ptrs && scratch && (numel == 0 || row_data), "NULL pointer");
  if (n_cols == 0 || numel == 0) return TCHGEO_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 4, stream));
  const int64_t grid = (n_cols + 255) / 256;
  TCHGEO_REQUIRE(grid < ((int64_t)1 << 31), "too many columns");
  cumsum_kernel<double><<<(unsigned)grid, 256, 0, stream>>>(col_ptrs, n_cols, row_data, numel, (uint32_t*)scratch);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  uint32_t h = 0;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(&h, scratch, 4, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  return status_from_dev_err(h);
}

static cudaError_t sort_edges_impl(void* tmp, size_t& tmp_bytes, const double* keys_in, double* keys_out,
                                   const int64_t* vals_in, int64_t* vals_out, int64_t numel, int64_t n_cols,
                                   const int64_t* col_ptrs, bool descending, cudaStream_t stream) {
  if (descending)
    return cub::DeviceSegmentedSort::StableSortPairsDescending(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out,
                                                               (int)numel, (int)n_cols, col_ptrs, col_ptrs + 1, stream);
  return cub::DeviceSegmentedSort::StableSortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int)numel,
                                                   (int)n_cols, col_ptrs, col_ptrs + 1, stream);
}

extern "C" size_t tchgeo_csc_sort_edges_workspace_bytes(int64_t numel, int64_t n_cols) {
  if (numel <= 0 || n_cols <= 0 || numel >= ((int64_t)1 << 31) || n_cols >= ((int64_t)1 << 31)) return 256;
  size_t tmp = 0;
  if (sort_edges_impl(nullptr, tmp, nullptr, nullptr, nullptr, nullptr, numel, n_cols, nullptr, false, 0) != cudaSuccess)
    return 0;
  return 256 + up256((size_t)numel * 8) + up256((size_t)(n_cols + 1) * 8) + up256(tmp);
}

extern "C" tchgeo_status tchgeo_csc_sort_edges(const int64_t* col_ptrs, int64_t n_cols, const int64_t* perm,
                                               const double* row_weights, int64_t numel, int32_t descending,
                                               int64_t* new_perm, void* workspace, size_t workspace_bytes,
                                               tchgeo_stream stream_) {
  TCHGEO_REQUIRE(n_cols >= 0 && numel >= 0, "negative size");
  TCHGEO_REQUIRE(numel < ((int64_t)1 << 31) && n_cols < ((int64_t)1 << 31), "too many edges or columns for one call");
  TCHGEO_REQUIRE(col_ptrs && (numel == 0 || (perm && row_weights && new_perm)), "NULL pointer");
  if (numel == 0) return TCHGEO_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  // columns with <= 1 element and positions outside every column keep their entry (new_perm = perm.copy(), :15)
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(new_perm, perm, (size_t)numel * 8, cudaMemcpyDeviceToDevice, stream));
  if (n_cols == 0) return TCHGEO_OK;
  const size_t need = tchgeo_csc_sort_edges_workspace_bytes(numel, n_cols);
  TCHGEO_REQUIRE(need != 0, "cub workspace query failed");
  if (!workspace || workspace_bytes < need) {
    set_last_error("workspace too small: %zu < %zu", workspace_bytes, need);
    return TCHGEO_ERR_CAPACITY;
  }
  double* keys_out = (double*)((char*)workspace + 256);
  int64_t* ptrs = (int64_t*)((char*)keys_out + up256((size_t)numel * 8));
  void* tmp = (char*)ptrs + up256((size_t)(n_cols + 1) * 8);
  size_t tmp_bytes = need - 256 - up256((size_t)numel * 8) - up256((size_t)(n_cols + 1) * 8);
  clamp_ptrs_kernel<<<(unsigned)((n_cols + 1 + 255) / 256), 256, 0, stream>>>(col_ptrs, n_cols + 1, numel, ptrs);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  TCHGEO_CUDA_CHECK(sort_edges_impl(tmp, tmp_bytes, row_weights, keys_out, perm, new_perm, numel, n_cols, ptrs,
                                    descending != 0, stream));
  return TCHGEO_OK;
}
