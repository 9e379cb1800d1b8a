// The graph handle of include/tchgeo_cuda.h (tchgeo_graph_create ... tchgeo_graph_destroy): borrowed CSC/CSR arrays
// of R relations plus the derived arrays the fast paths read.  Implemented in neighbor_sampling.cu.
#pragma once
#include <vector>

#include "common.cuh"

struct tchgeo_graph {
  int R = 0;
  int device = 0;
  std::vector<const int64_t*> ptrs, indices;
  std::vector<int64_t> num_major, nnz;
  std::vector<const double*> weights;
  std::vector<const int64_t*> timestamps;
  // derived, owned
  std::vector<int32_t*> indices32;
  std::vector<uint8_t> replica_state;  // 0 = not tried, 1 = built, 2 = not representable (ids >= 2^31) or disabled
  std::vector<int64_t> max_index;      // largest entry of indices[r], found while the replica is built (-1 = unknown)
  std::vector<double2*> wrec;
  size_t derived_bytes = 0;
};

namespace tchgeo {
// builds the derived arrays named by `what` (TCHGEO_PREPARE_* bits) that are still missing; synchronises when it builds
tchgeo_status graph_ensure(tchgeo_graph* g, int32_t what, cudaStream_t stream);
}  // namespace tchgeo
