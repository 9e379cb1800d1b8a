// Frontier dedup + insertion-order local-id relabeling of a sampled tree (additive stage).
//
// Semantic of src/algo/negative_sampling.rs:20-47 (samples_mapping):
//   nodes     = seeds (all, duplicates kept, :25) ++ every non-seed id at its first appearance (:36-39)
//   map[seed] = index of the seed's LAST occurrence (HashMap::extend overwrites, :26)
//   local[i]  = map[samples[i]]
// Parallel formulation: a global-memory open-addressing hash insert resolves, per distinct id,
// max(seed position) and min(non-seed position); "first occurrence" flags are then compacted by a
// stable exclusive scan, which reproduces the serial insertion order exactly.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace tchgeo {
namespace {

constexpr int RL_THREADS = 256;
constexpr int64_t RL_EMPTY = -1;  // node ids are non-negative; 0xFF.. memset initialises everything

__device__ __forceinline__ uint64_t rl_hash(int64_t key) {
  uint64_t h = (uint64_t)key * 0x9E3779B97F4A7C15ull;
  return h ^ (h >> 29);
}

__global__ void __launch_bounds__(RL_THREADS) rl_insert_kernel(const int64_t* __restrict__ samples, int64_t n,
                                                              int64_t num_seeds, unsigned long long* keys,
                                                              int* seed_last, unsigned* min_pos, uint64_t mask,
                                                              int* __restrict__ slot_of, uint32_t* err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t key = samples[i];
  if (key < 0) {
    atomicOr(err, DEV_ERR_INDEX);
    slot_of[i] = -1;
    return;
  }
  uint64_t h = rl_hash(key) & mask;
  while (true) {
    const unsigned long long old = atomicCAS(keys + h, (unsigned long long)RL_EMPTY, (unsigned long long)key);
    if (old == (unsigned long long)RL_EMPTY || old == (unsigned long long)key) break;
    h = (h + 1) & mask;
  }
  slot_of[i] = (int)h;
  if (i < num_seeds) atomicMax(seed_last + h, (int)i);
  else atomicMin(min_pos + h, (unsigned)i);
}

__global__ void __launch_bounds__(RL_THREADS) rl_flag_kernel(int64_t n, int64_t num_seeds, const int* __restrict__ seed_last,
                                                            const unsigned* __restrict__ min_pos,
                                                            const int* __restrict__ slot_of, int* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int f = 0;
  const int h = slot_of[i];
  if (i < num_seeds) f = 1;
  else if (h >= 0) f = (seed_last[h] < 0 && min_pos[h] == (unsigned)i) ? 1 : 0;
  flags[i] = f;
}

__global__ void __launch_bounds__(RL_THREADS) rl_assign_kernel(const int64_t* __restrict__ samples, int64_t n,
                                                              int64_t num_seeds, const int* __restrict__ seed_last,
                                                              const int* __restrict__ slot_of,
                                                              const int* __restrict__ flags, const int* __restrict__ ranks,
                                                              int* val, int64_t* __restrict__ nodes, int64_t* total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flags[i]) {
    const int r = ranks[i];
    nodes[r] = samples[i];
    const int h = slot_of[i];
    if (h >= 0 && (i >= num_seeds || seed_last[h] == (int)i)) val[h] = r;
  }
  if (i == n - 1) *total = (int64_t)ranks[i] + flags[i];
}

__global__ void __launch_bounds__(RL_THREADS) rl_lookup_kernel(int64_t n, const int* __restrict__ slot_of,
                                                              const int* __restrict__ val, int64_t* __restrict__ local) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int h = slot_of[i];
  local[i] = h >= 0 ? (int64_t)val[h] : -1;
}

struct RlLayout {
  uint64_t cap;
  size_t off_keys, off_seed_last, off_min_pos, off_val, off_slot_of, off_flags, off_ranks, off_total, off_cub, cub_bytes, total;
};
inline size_t rl_align(size_t x) { return (x + 255) / 256 * 256; }

cudaError_t rl_layout(int64_t n, RlLayout& L) {
  uint64_t cap = 1024;
  while (cap < 2ull * (uint64_t)n + 2) cap <<= 1;
  L.cap = cap;
  size_t cub_bytes = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const int*)nullptr, (int*)nullptr,
                                                (int64_t)(n > 0 ? n : 1));
  if (e != cudaSuccess) return e;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  size_t o = 0;
  L.off_total = o; o += 256;       // [0] total (i64), [8] err (u32)
  L.off_keys = o; o += rl_align(cap * 8);
  L.off_seed_last = o; o += rl_align(cap * 4);
  L.off_min_pos = o; o += rl_align(cap * 4);
  L.off_val = o; o += rl_align(cap * 4);
  L.off_slot_of = o; o += rl_align(nn * 4);
  L.off_flags = o; o += rl_align(nn * 4);
  L.off_ranks = o; o += rl_align(nn * 4);
  L.off_cub = o; o += rl_align(cub_bytes);
  L.cub_bytes = cub_bytes;
  L.total = o + 256;
  return cudaSuccess;
}

}  // namespace
}  // namespace tchgeo

using namespace tchgeo;

extern "C" size_t tchgeo_unique_relabel_workspace_bytes(int64_t n) {
  if (n < 0 || n >= ((int64_t)1 << 30)) return 0;
  RlLayout L;
  if (rl_layout(n, L) != cudaSuccess) return 0;
  return L.total;
}

extern "C" tchgeo_status tchgeo_unique_relabel(const int64_t* samples, int64_t n, int64_t num_seeds, int64_t* nodes,
                                               int64_t* local, int64_t* num_nodes, void* workspace,
                                               size_t workspace_bytes, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(n >= 0 && n < ((int64_t)1 << 30), "n out of range");
  TCHGEO_REQUIRE(num_seeds >= 0 && num_seeds <= n, "num_seeds out of range");
  if (n == 0) {
    if (num_nodes) *num_nodes = 0;
    return TCHGEO_OK;
  }
  TCHGEO_REQUIRE(samples && nodes && local && workspace, "NULL pointer");
  RlLayout L;
  TCHGEO_CUDA_CHECK(rl_layout(n, L));
  if (workspace_bytes < L.total) {
    set_last_error("workspace too small: need %zu bytes, got %zu", L.total, workspace_bytes);
    return TCHGEO_ERR_CAPACITY;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  int64_t* d_total = (int64_t*)(ws + L.off_total);
  uint32_t* d_err = (uint32_t*)(ws + L.off_total + 8);
  unsigned long long* keys = (unsigned long long*)(ws + L.off_keys);
  int* seed_last = (int*)(ws + L.off_seed_last);
  unsigned* min_pos = (unsigned*)(ws + L.off_min_pos);
  int* val = (int*)(ws + L.off_val);
  int* slot_of = (int*)(ws + L.off_slot_of);
  int* flags = (int*)(ws + L.off_flags);
  int* ranks = (int*)(ws + L.off_ranks);
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(ws + L.off_total, 0, 256, stream));
  // keys = -1 (empty), seed_last = -1, min_pos = 0xFFFFFFFF: one 0xFF memset over the three tables
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(ws + L.off_keys, 0xFF, L.off_val - L.off_keys, stream));
  const unsigned grid = (unsigned)((n + RL_THREADS - 1) / RL_THREADS);
  rl_insert_kernel<<<grid, RL_THREADS, 0, stream>>>(samples, n, num_seeds, keys, seed_last, min_pos, L.cap - 1, slot_of,
                                                    d_err);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  rl_flag_kernel<<<grid, RL_THREADS, 0, stream>>>(n, num_seeds, seed_last, min_pos, slot_of, flags);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  size_t cub_bytes = L.cub_bytes;
  TCHGEO_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ws + L.off_cub, cub_bytes, (const int*)flags, ranks, n, stream));
  rl_assign_kernel<<<grid, RL_THREADS, 0, stream>>>(samples, n, num_seeds, seed_last, slot_of, flags, ranks, val, nodes,
                                                    d_total);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  rl_lookup_kernel<<<grid, RL_THREADS, 0, stream>>>(n, slot_of, val, local);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  int64_t h[2] = {0, 0};
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(h, d_total, 16, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (num_nodes) *num_nodes = h[0];
  return status_from_dev_err((uint32_t)h[1]);
}
