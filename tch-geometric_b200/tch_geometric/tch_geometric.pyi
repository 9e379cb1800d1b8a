# Type stubs of the B200 path; signatures are those of the reference's tch_geometric/tch_geometric.pyi:12-44
# and :83-91 (trailing Optional arguments are optional, as in pyo3 0.15).
from typing import Dict, List, Optional, Tuple, Union

from torch import Tensor

from tch_geometric.utils import EdgeFilter, EdgeSampler

NodeType = str
EdgeType = Tuple[str, str, str]
LayerOffset = Tuple[int, int, int]
RelType = str

def to_csc(row_col: Tensor, size: Union[int, Tuple[int, int]]) -> Tuple[Tensor, Tensor, Tensor]: ...
def to_csr(row_col: Tensor, size: Union[int, Tuple[int, int]]) -> Tuple[Tensor, Tensor, Tensor]: ...
def neighbor_sampling_homogenous(
    col_ptrs: Tensor,
    row_indices: Tensor,
    inputs: Tensor,
    num_neighbors: List[int],
    sampler: Optional[EdgeSampler] = ...,
    filter: Optional[Tuple[EdgeFilter, Tensor]] = ...,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, List[LayerOffset]]: ...
def neighbor_sampling_heterogenous(
    node_types: List[NodeType],
    edge_types: List[EdgeType],
    col_ptrs: Dict[RelType, Tensor],
    row_indices: Dict[RelType, Tensor],
    inputs: Dict[NodeType, Tensor],
    num_neighbors: Dict[RelType, List[int]],
    num_hops: int,
    sampler: Optional[EdgeSampler] = ...,
    filter: Optional[Tuple[EdgeFilter, Tensor]] = ...,
) -> Tuple[
    Dict[NodeType, Tensor], Dict[RelType, Tensor], Dict[RelType, Tensor], Dict[RelType, Tensor],
    Dict[RelType, List[LayerOffset]],
]: ...
def random_walk(
    row_ptrs: Tensor, col_indices: Tensor, start: Tensor, walk_length: int, p: float, q: float
) -> Tensor: ...
# SURVEY 8(f) rows built so far; signatures of the reference's .pyi:94-105 and :122-146
def tempo_random_walk(
    row_ptrs: Tensor, col_indices: Tensor, node_timestamps: Tensor, edge_timestamps: Tensor, start: Tensor,
    start_timestamps: Tensor, walk_length: int, window: Tuple[int, int],
) -> Tuple[Tensor, Tensor]: ...
def negative_sample_neighbors_homogenous(
    row_ptrs: Tensor, col_indices: Tensor, graph_size: Tuple[int, int], inputs: Tensor, num_neg: int, try_count: int,
) -> Tuple[Tensor, Tensor, Tensor, int]: ...
def negative_sample_neighbors_heterogenous(
    node_types: List[NodeType],
    edge_types: List[EdgeType],
    row_ptrs: Dict[RelType, Tensor],
    col_indices: Dict[RelType, Tensor],
    sizes: Dict[RelType, Tuple[int, int]],
    inputs: Dict[NodeType, Tensor],
    num_neg: int,
    try_count: int,
    inbound: bool,
) -> Tuple[Dict[NodeType, Tensor], Dict[RelType, Tensor], Dict[RelType, Tensor], Dict[NodeType, int]]: ...
# src/data/transform.rs (not exported by the reference's Python module) and the gather that follows the sampler
def csc_edge_cumsum(col_ptrs: Tensor, row_data: Tensor) -> None: ...
def csc_sort_edges(col_ptrs: Tensor, perm: Tensor, row_weights: Tensor, descending: bool = ...) -> Tensor: ...
def gather_rows(src: Tensor, index: Tensor) -> Tensor: ...
