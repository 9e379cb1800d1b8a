"""tch_geometric -- B200-native (sm_100a) drop-in for the reference module's sampling hot path.

Mirrors `tch_geometric/__init__.py:1-2` + `tch_geometric.pyi` of the reference for the functions on
the path (to_csc, to_csr, neighbor_sampling_homogenous, neighbor_sampling_heterogenous, random_walk) and the
"next" rows of SURVEY 8(f) built so far (negative_sample_neighbors_*, csc_edge_cumsum, csc_sort_edges);
everything else in the reference (hgt/budget sampling, biased temporal walks) stays on the reference's CPU
implementation and is not provided here.  Importing this package loads
libtchgeo_cuda.so and fails if it has not been built: there is no CPU fallback.
"""
from . import _native  # noqa: F401  (loads the CUDA library, raises if missing)
from .ops import (  # noqa: F401
    GraphHandle,
    HeterogenousSampler,
    HomogenousSampler,
    HostBatches,
    SampledBatches,
    clear_caches,
    csc_edge_cumsum,
    csc_sort_edges,
    gather_rows,
    ind2ptr,
    negative_sample_neighbors_heterogenous,
    negative_sample_neighbors_homogenous,
    neighbor_sampling_heterogenous,
    neighbor_sampling_homogenous,
    neighbor_sampling_homogenous_batched,
    random_walk,
    rel_key,
    rng_reseed,
    set_l2_fetch_granularity,
    tempo_random_walk,
    to_csc,
    to_csr,
    unique_relabel,
    unique_relabel_batched,
)
from .utils import (  # noqa: F401
    TEMPORAL_SAMPLE_DYNAMIC,
    TEMPORAL_SAMPLE_RELATIVE,
    TEMPORAL_SAMPLE_STATIC,
    EdgeFilter,
    EdgeSampler,
    TemporalEdgeFilter,
    UniformEdgeSampler,
    WeightedEdgeSampler,
)
from ._native import ReferencePanic  # noqa: F401
