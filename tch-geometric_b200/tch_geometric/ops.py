"""Host side of the B200 sampling path: the Python mirror of the reference's binding layer
(src/python.rs) for to_csc / to_csr / neighbor_sampling_homogenous / neighbor_sampling_heterogenous /
random_walk.  Same names, argument meaning, return layouts and error behaviour as
tch_geometric/tch_geometric.pyi; tensors live on a CUDA device instead of the CPU
(src/utils/tensor.rs:50-52 becomes a CUDA-device check on this path).

torch is used only for device memory and streams; all compute happens behind the C ABI of
libtchgeo_cuda.so (include/tchgeo_cuda.h).  There is no CPU fallback.
"""
import ctypes
import os
import threading
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from . import _native as N

LayerOffset = Tuple[int, int, int]

# ---------------------------------------------------------------------------------------------
# process-global RNG, mirroring src/utils/random.rs:8-23: entropy-seeded, every API call forks a
# child stream (here: a fresh 64-bit Philox key).  rng_reseed exists in the reference but has no
# Python binding (python.rs:785-796); it is exported here because tests need reproducibility.
# ---------------------------------------------------------------------------------------------
_rng_lock = threading.Lock()
_rng_state = int.from_bytes(os.urandom(8), "little")
_MASK64 = (1 << 64) - 1


def rng_reseed(seed: int) -> None:
    global _rng_state
    with _rng_lock:
        _rng_state = int(seed) & _MASK64


def splitmix64(state: int) -> Tuple[int, int]:
    """One splitmix64 step: (state) -> (new_state, output)."""
    state = (state + 0x9E3779B97F4A7C15) & _MASK64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
    return state, z ^ (z >> 31)


def _rng_get() -> int:
    """One child seed per API call (random.rs:19-23)."""
    global _rng_state
    with _rng_lock:
        _rng_state, out = splitmix64(_rng_state)
    return out


# ---------------------------------------------------------------------------------------------
# tensor plumbing (src/utils/tensor.rs:10-70)
# ---------------------------------------------------------------------------------------------
def _check(t, dtype, name, device=None):
    if not isinstance(t, Tensor):
        raise ValueError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"Tensor must be on Cuda device ({name} is on {t.device})")  # InvalidDevice
    if device is not None and t.device != device:
        raise ValueError(f"Tensor must be on {device} device ({name} is on {t.device})")
    if t.dtype != dtype:
        raise ValueError(f"Tensor must be a is of invalid type. Expected {dtype} but got {t.dtype} ({name})")
    if not t.is_contiguous():
        # the reference silently reads garbage here (quirk Q10); refuse instead
        raise ValueError(f"{name} must be contiguous")
    return t


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def _size_tuple(size) -> Tuple[int, int]:
    """GraphSize, python.rs:12-25"""
    if isinstance(size, (tuple, list)):
        if len(size) != 2:
            raise ValueError("size must be an int or a pair of ints")
        return int(size[0]), int(size[1])
    return int(size), int(size)


def set_l2_fetch_granularity(nbytes: int = 32, device=None) -> int:
    """Optional device hint (extension): L2->DRAM fetch granularity for the gather-bound sampling path.
    Returns the value the driver reports afterwards."""
    actual = ctypes.c_int32(0)
    with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
        N.check(N.lib.tchgeo_device_set_l2_fetch_granularity(int(nbytes), ctypes.addressof(actual)))
    return actual.value


# ---------------------------------------------------------------------------------------------
# to_csc / to_csr (python.rs:27-53 -> storage.rs:103-127)
# ---------------------------------------------------------------------------------------------
def _to_csx(row_col: Tensor, size, csc: bool):
    _check(row_col, torch.int64, "row_col")
    if row_col.dim() != 2 or row_col.shape[0] != 2:
        raise ValueError("Tensor must be of rank [2, E]")
    n_rows, n_cols = _size_tuple(size)
    dev = row_col.device
    E = row_col.shape[1]
    n_major = n_cols if csc else n_rows
    with torch.cuda.device(dev):
        ptrs = torch.empty(n_major + 1, dtype=torch.int64, device=dev)
        indices = torch.empty(E, dtype=torch.int64, device=dev)
        perm = torch.empty(E, dtype=torch.int64, device=dev)
        ws_bytes = N.lib.tchgeo_coo_to_csx_workspace_bytes(E, n_rows, n_cols)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        st = N.lib.tchgeo_coo_to_csx(_ptr(row_col[0]), _ptr(row_col[1]), E, n_rows, n_cols, 1 if csc else 0,
                                     _ptr(ptrs), _ptr(indices), _ptr(perm), _ptr(ws), ws_bytes, _stream(dev))
    N.check(st)
    return ptrs, indices, perm


def to_csc(row_col: Tensor, size: Union[int, Tuple[int, int]]) -> Tuple[Tensor, Tensor, Tensor]:
    return _to_csx(row_col, size, True)


def to_csr(row_col: Tensor, size: Union[int, Tuple[int, int]]) -> Tuple[Tensor, Tensor, Tensor]:
    return _to_csx(row_col, size, False)


def ind2ptr(ind: Tensor, m: int) -> Tensor:
    """storage.rs:67-101 (not exported by the reference's Python module; exposed for its KAT)."""
    _check(ind, torch.int64, "ind")
    out = torch.empty(m + 1, dtype=torch.int64, device=ind.device)
    with torch.cuda.device(ind.device):
        N.check(N.lib.tchgeo_ind2ptr(_ptr(ind), ind.numel(), m, _ptr(out), _stream(ind.device)))
    return out


# ---------------------------------------------------------------------------------------------
# sampler / filter argument extraction (python.rs:107-168)
# ---------------------------------------------------------------------------------------------
def _extract_sampler(sampler, hetero: bool):
    """SamplerType is tried in declaration order: Uniform(with_replacement) then Weighted(weights)."""
    if sampler is None:
        return N.SAMPLER_UNIFORM, None
    wr = getattr(sampler, "with_replacement", None)
    if isinstance(wr, bool):
        return (N.SAMPLER_UNIFORM_REPLACE if wr else N.SAMPLER_UNIFORM), None
    w = getattr(sampler, "weights", None)
    if w is not None:
        if hetero != isinstance(w, dict):
            raise ValueError("Unknown error: data must be %s" % ("heterogenous" if hetero else "homogenous"))
        return N.SAMPLER_WEIGHTED, w
    # pyo3 fails the Option<SamplerType> extraction with a TypeError
    raise TypeError("sampler must have a `with_replacement: bool` or a `weights` attribute")


def _extract_filter(filter, hetero: bool):
    """FilterType::TemporalFilter((TemporalFilter{window, timestamps, forward, mode}, inputs_state)), python.rs:137-168.
    -> None | (abi_mode, forward, (lo, hi), timestamps, inputs_state).  As in python.rs:219-249 a mode outside
    {STATIC, RELATIVE, DYNAMIC} falls through to the IdentityFilter arm."""
    if filter is None:
        return None
    try:
        ft, inputs_state = filter
        window, timestamps = ft.window, ft.timestamps
        forward, mode = bool(ft.forward), int(ft.mode)
        lo, hi = int(window[0]), int(window[1])
    except (TypeError, ValueError, AttributeError) as exc:
        raise TypeError("filter must be a (TemporalEdgeFilter, inputs_state) pair") from exc
    for name, v in (("timestamps", timestamps), ("inputs_state", inputs_state)):
        if hetero != isinstance(v, dict):
            raise ValueError("Unknown error: data must be %s" % ("heterogenous" if hetero else "homogenous"))
    if mode not in (0, 1, 2):
        return None
    return mode + 1, forward, (lo, hi), timestamps, inputs_state


# ---------------------------------------------------------------------------------------------
# graph handles (tchgeo_graph_t).  The derived arrays of the fast paths -- the int32 replica of the indices (random
# gathers span half as many DRAM lines; hop 3 on B200: DRAM reads 6.28 -> 4.86 GB, 2.24 -> 1.88 ms) and the
# (weight, prefix sum) records of the weighted sampler -- are built and owned by the library behind the handle.  The
# reference API is stateless (tensors are passed on every call), so this mirror keeps one handle per set of graph
# tensors, found again by storage address + version counter.  TCHGEO_INDEX_REPLICA=0 / TCHGEO_WEIGHT_CUMSUM=0 (read by
# the library) disable the derived arrays.
# ---------------------------------------------------------------------------------------------
_GRAPH_CACHE_MAX = 16


class GraphHandle:
    """Owns a tchgeo_graph_t over R relations' (ptrs, indices[, weights]) tensors and keeps those tensors alive: the
    memory cannot be freed and handed to a different tensor at the same address while the handle lives."""

    def __init__(self, ptrs, indices, weights=None):
        R = len(ptrs)
        self._tensors = (list(ptrs), list(indices), list(weights) if weights is not None else None)
        tab = lambda ts: np.array([(t.data_ptr() if t is not None else 0) for t in ts], dtype=np.uint64)
        num_major = np.array([(max(t.numel() - 1, 0) if t is not None else 0) for t in ptrs], dtype=np.int64)
        nnz = np.array([(t.numel() if t is not None else 0) for t in indices], dtype=np.int64)
        h = ctypes.c_void_p(0)
        p_tab, i_tab = tab(ptrs), tab(indices)
        N.check(N.lib.tchgeo_graph_create(R, p_tab.ctypes.data, num_major.ctypes.data, i_tab.ctypes.data, nnz.ctypes.data,
                                          ctypes.addressof(h)))
        self.handle, self.num_rels = h, R
        if weights is not None:
            w_tab = tab(weights)
            N.check(N.lib.tchgeo_graph_set_weights(self.handle, w_tab.ctypes.data))

    def prepare(self, what, device):
        with torch.cuda.device(device):
            N.check(N.lib.tchgeo_graph_prepare(self.handle, int(what), _stream(device)))

    @property
    def derived_bytes(self):
        return int(N.lib.tchgeo_graph_derived_bytes(self.handle))

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                N.lib.tchgeo_graph_destroy(h)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass


class _GraphCache:
    """LRU of GraphHandles keyed by their source tensors' storage; an in-place edit bumps `_version` and misses."""

    def __init__(self):
        self._d = {}

    @staticmethod
    def _key(groups):
        return tuple(None if g is None else tuple(None if t is None else (t.data_ptr(), t.numel(), t.device.index, t._version)
                                                  for t in g) for g in groups)

    def get(self, ptrs, indices, weights=None):
        k = self._key((ptrs, indices, weights))
        hit = self._d.pop(k, None)
        if hit is None:
            while len(self._d) >= _GRAPH_CACHE_MAX:
                self._d.pop(next(iter(self._d)))
            hit = GraphHandle(ptrs, indices, weights)
        self._d[k] = hit  # most recently used last
        return hit

    def clear(self):
        self._d.clear()


_graph_cache = _GraphCache()


def clear_caches() -> None:
    """Drop the cached graph handles: frees their derived device arrays and releases the source tensors.  Call it when
    switching graphs (a handle pins its tensors, so a dropped graph is not freed while its handle is cached)."""
    _graph_cache.clear()


# ---------------------------------------------------------------------------------------------
# per-column transforms of CSC edge data (src/data/transform.rs; not exported by the reference's Python module,
# exposed here because the weighted sampler uses the prefix sums and the reference carries KATs for both)
# ---------------------------------------------------------------------------------------------
def csc_edge_cumsum(col_ptrs: Tensor, row_data: Tensor) -> None:
    """transform.rs:36-60: in-place inclusive prefix sum of `row_data` (f64) inside every column, serial order."""
    _check(col_ptrs, torch.int64, "col_ptrs")
    dev = col_ptrs.device
    _check(row_data, torch.float64, "row_data", dev)
    scratch = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.tchgeo_csc_edge_cumsum_f64(_ptr(col_ptrs), max(col_ptrs.numel() - 1, 0), _ptr(row_data),
                                                 row_data.numel(), _ptr(scratch), _stream(dev)))


def csc_sort_edges(col_ptrs: Tensor, perm: Tensor, row_weights: Tensor, descending: bool = False) -> Tensor:
    """transform.rs:7-34: new_perm = perm with every column's entries reordered by their weight (stable)."""
    _check(col_ptrs, torch.int64, "col_ptrs")
    dev = col_ptrs.device
    _check(perm, torch.int64, "perm", dev)
    _check(row_weights, torch.float64, "row_weights", dev)
    if perm.numel() != row_weights.numel():
        raise ValueError("perm and row_weights must have the same length")
    n, ncols = perm.numel(), max(col_ptrs.numel() - 1, 0)
    out = torch.empty_like(perm)
    ws_bytes = N.lib.tchgeo_csc_sort_edges_workspace_bytes(n, ncols)
    ws = torch.empty(max(int(ws_bytes), 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.tchgeo_csc_sort_edges(_ptr(col_ptrs), ncols, _ptr(perm), _ptr(row_weights), n,
                                            1 if descending else 0, _ptr(out), _ptr(ws), ws_bytes, _stream(dev)))
    return out


# ---------------------------------------------------------------------------------------------
# generic driver over tchgeo_neighbor_sampling
# ---------------------------------------------------------------------------------------------
class _Call:
    """One tchgeo_plan_t: the argument tables, the output tensors and the workspace it borrows.  The plan handle keeps a
    deep copy of the arguments, the launch plan and the pinned length table behind the ABI; this class only allocates
    device memory through torch and hands pointers over."""

    def __init__(self, device, rel_src, rel_dst, col_ptrs, row_indices, weights, fanouts, rel_active, inputs, seeds,
                 num_batches, num_hops, sampler_kind, seed, batch_base=0, filt=None, relabel=False):
        """filt: None | (abi_mode, forward, (lo, hi), [timestamps per relation], [inputs_state per node type])"""
        T, R, H, B = len(seeds), len(rel_src), num_hops, num_batches
        self.T, self.R, self.H, self.B, self.device = T, R, H, B, device
        self.plan = None
        self.max_fanout = int(np.max(np.asarray(fanouts, dtype=np.int64), initial=0))
        # the reference slices weights / timestamps with the column's range and panics when they are short
        # (EdgeAttr::get, src/data/graph.rs:103-120); the kernels check every column against nnz as well
        for r in range(R):
            nnz = row_indices[r].numel() if row_indices[r] is not None else 0
            for name, ts in (("weights", weights), ("timestamps", filt[3] if filt is not None else None)):
                if ts is not None and ts[r] is not None and ts[r].numel() < nnz:
                    raise N.ReferencePanic(f"{name} of relation {r} has {ts[r].numel()} entries, row_indices {nnz}")
        a = N.SamplingArgs()
        self.args = a
        k = self._keep = []

        def host(arr, dtype):
            x = np.ascontiguousarray(np.asarray(arr, dtype=dtype))
            k.append(x)
            return x

        def ptr_table(tensors):
            x = np.array([(t.data_ptr() if t is not None else 0) for t in tensors], dtype=np.uint64)
            k.append(x)
            k.append(list(tensors))
            return x

        a.num_node_types, a.num_rels, a.num_hops, a.sampler_kind = T, R, H, sampler_kind
        a.rel_src = host(rel_src, np.int32).ctypes.data
        a.rel_dst = host(rel_dst, np.int32).ctypes.data
        self.graph = _graph_cache.get(col_ptrs, row_indices, weights)
        a.graph = self.graph.handle
        a.fanouts = host(fanouts, np.int64).ctypes.data
        a.rel_active = host(rel_active, np.uint8).ctypes.data
        a.num_batches = B
        a.inputs = ptr_table(inputs).ctypes.data
        a.seeds_per_batch = host(seeds, np.int64).ctypes.data
        a.seed = seed
        a.batch_base = batch_base
        cap_n = np.zeros(T, dtype=np.int64)
        cap_e = np.zeros(R, dtype=np.int64)
        N.check(N.lib.tchgeo_neighbor_sampling_capacity(ctypes.byref(a), cap_n.ctypes.data, cap_e.ctypes.data))
        self.cap_n, self.cap_e = cap_n, cap_e
        k += [cap_n, cap_e]
        i64 = dict(dtype=torch.int64, device=device)
        self.samples = [torch.empty((B, int(c)), **i64) for c in cap_n]
        self.rows = [torch.empty((B, int(c)), **i64) for c in cap_e]
        self.cols = [torch.empty((B, int(c)), **i64) for c in cap_e]
        self.eidx = [torch.empty((B, int(c)), **i64) for c in cap_e]
        a.samples = ptr_table(self.samples).ctypes.data
        a.samples_stride = cap_n.ctypes.data
        a.rows = ptr_table(self.rows).ctypes.data
        a.cols = ptr_table(self.cols).ctypes.data
        a.edge_index = ptr_table(self.eidx).ctypes.data
        a.edges_stride = cap_e.ctypes.data
        self.nodes = self.local = None
        if relabel:
            self.nodes = [torch.empty((B, int(c)), **i64) for c in cap_n]
            self.local = [torch.empty((B, int(c)), **i64) for c in cap_n]
            a.nodes = ptr_table(self.nodes).ctypes.data
            a.local = ptr_table(self.local).ctypes.data
        self.states = None
        if filt is not None:
            a.filter_mode, a.filter_forward = int(filt[0]), int(filt[1])
            a.filter_window_lo, a.filter_window_hi = filt[2]
            a.timestamps = ptr_table(filt[3]).ctypes.data
            a.inputs_state = ptr_table(filt[4]).ctypes.data
            self.states = [torch.empty((B, int(c)), **i64) for c in cap_n]
            a.states = ptr_table(self.states).ctypes.data
        with torch.cuda.device(device):
            a.stream = torch.cuda.current_stream(device).cuda_stream
            ws_bytes = N.lib.tchgeo_neighbor_sampling_workspace_bytes(ctypes.byref(a))  # builds the derived arrays
            if ws_bytes == 0:
                raise ValueError(N.last_error() or "cannot plan this call")
            self.workspace = torch.empty(int(ws_bytes), dtype=torch.uint8, device=device)
            a.workspace = self.workspace.data_ptr()
            a.workspace_bytes = ws_bytes
            h = ctypes.c_void_p(0)
            N.check(N.lib.tchgeo_plan_create(ctypes.byref(a), ctypes.addressof(h)))
        self.plan = h
        self.num_launches = int(N.lib.tchgeo_plan_num_launches(h))
        # results live in host arrays owned by the plan (rewritten by every collect)
        ptrs = [ctypes.POINTER(ctypes.c_int64)() for _ in range(4)]
        N.check(N.lib.tchgeo_plan_results(h, *[ctypes.byref(x) for x in ptrs]))
        view = lambda ptr, shape: np.ctypeslib.as_array(ptr, shape=shape)
        self.samples_len = view(ptrs[0], (B, T))
        self.edges_len = view(ptrs[1], (B, R))
        self.layer_offsets = view(ptrs[2], (B, R, max(H, 1), 3))
        self.nodes_len = view(ptrs[3], (B, T)) if relabel else None

    def __del__(self):
        h, self.plan = getattr(self, "plan", None), None
        if h:
            try:
                N.lib.tchgeo_plan_destroy(h)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass

    def run(self, seed=None, batch_base=None, timed=False, collect=True):
        """collect=False only enqueues the launches on the current stream (no host synchronisation); `collect()`
        waits for them and decodes lengths and layer offsets.  Nothing else may use this call's buffers in between."""
        a = self.args
        if seed is not None:
            a.seed = seed
        if batch_base is not None:
            a.batch_base = batch_base
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            if timed:
                ms = np.zeros(max(self.R * max(self.H, 1) + 1, 1), dtype=np.float32)
                n = ctypes.c_int32(0)
                N.check(N.lib.tchgeo_plan_enqueue_timed(self.plan, a.seed, a.batch_base, stream, ms.ctypes.data, ms.size,
                                                        ctypes.addressof(n)))
                self.launch_ms = ms[:n.value].copy()
                return self
            N.check(N.lib.tchgeo_plan_enqueue(self.plan, a.seed, a.batch_base, stream))
        return self.collect() if collect else self

    def collect(self):
        """Second half of run(collect=False): waits for the event recorded behind the step's launches (not for the whole
        stream) and turns the device-side error word into an exception."""
        N.check(N.lib.tchgeo_plan_collect(self.plan))
        return self


def _as_seed_matrix(inputs: Tensor, device, name="inputs") -> Tensor:
    if not isinstance(inputs, Tensor):
        raise ValueError(f"{name} must be a torch.Tensor")
    if inputs.dtype != torch.int64:
        raise ValueError(f"Tensor must be a is of invalid type. Expected torch.int64 but got {inputs.dtype} ({name})")
    if inputs.device != device:
        # seeds usually come from a host-side loader: one small H2D copy (async if pinned)
        inputs = inputs.to(device, non_blocking=True)
    return inputs.contiguous()


def neighbor_sampling_homogenous(
        col_ptrs: Tensor,
        row_indices: Tensor,
        inputs: Tensor,
        num_neighbors: List[int],
        sampler=None,
        filter=None,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, List[LayerOffset]]:
    """python.rs:187-271.  Returns (samples, rows, cols, edge_index, layer_offsets)."""
    call = _homogenous_call(col_ptrs, row_indices, inputs, num_neighbors, sampler, filter, batched=False)
    call.run(seed=_rng_get())
    ns, ne = int(call.samples_len[0, 0]), int(call.edges_len[0, 0])
    lo = [tuple(int(x) for x in call.layer_offsets[0, 0, h]) for h in range(call.H)]
    return call.samples[0][0, :ns], call.rows[0][0, :ne], call.cols[0][0, :ne], call.eidx[0][0, :ne], lo


def _homogenous_call(col_ptrs, row_indices, inputs, num_neighbors, sampler, filter, batched, relabel=False):
    flt = _extract_filter(filter, hetero=False)
    _check(col_ptrs, torch.int64, "col_ptrs")
    dev = col_ptrs.device
    _check(row_indices, torch.int64, "row_indices", dev)
    kind, w = _extract_sampler(sampler, hetero=False)
    if w is not None:
        _check(w, torch.float64, "weights", dev)
    inputs = _as_seed_matrix(inputs, dev)
    if batched:
        if inputs.dim() != 2:
            raise ValueError("batched inputs must have shape [B, S]")
        B, S = inputs.shape
    else:
        B, S = 1, inputs.numel()
    fan = [int(k) for k in num_neighbors]
    if any(k < 0 for k in fan):
        raise OverflowError("can't convert negative int to unsigned")  # Vec<usize> extraction
    filt = None
    if flt is not None:
        mode, forward, window, ts, st = flt
        _check(ts, torch.int64, "timestamps", dev)
        st = _as_seed_matrix(st, dev, "inputs_state")
        if st.numel() != inputs.numel():
            # the reference indexes states[i] for every frontier index and panics when it is short (quirk Q10)
            raise N.ReferencePanic("inputs_state must have one entry per input")
        filt = (mode, forward, window, [ts], [st])
    return _Call(dev, [0], [0], [col_ptrs], [row_indices], [w] if w is not None else None, fan, [1], [inputs], [S],
                 max(B, 1), len(fan), kind, 0, filt=filt, relabel=relabel)


class SampledBatches:
    """Result of neighbor_sampling_homogenous_batched: B independent reference-layout results that
    share padded [B, capacity] buffers.  `batch(b)` is exactly what the reference returns for inputs[b]."""

    def __init__(self, call):
        self._call = call
        self.samples, self.rows, self.cols, self.edge_index = call.samples[0], call.rows[0], call.cols[0], call.eidx[0]
        self.samples_len = call.samples_len[:, 0]
        self.edges_len = call.edges_len[:, 0]
        self.layer_offsets = call.layer_offsets[:, 0]
        # dedup + insertion-order relabel stage (K7), when the call asked for it
        self.nodes = call.nodes[0] if call.nodes is not None else None
        self.local = call.local[0] if call.local is not None else None
        self.nodes_len = call.nodes_len[:, 0] if call.nodes_len is not None else None

    def __len__(self):
        return self._call.B

    def batch(self, b):
        ns, ne = int(self.samples_len[b]), int(self.edges_len[b])
        lo = [tuple(int(x) for x in self.layer_offsets[b, h]) for h in range(self._call.H)]
        return self.samples[b, :ns], self.rows[b, :ne], self.cols[b, :ne], self.edge_index[b, :ne], lo

    def to_host(self, host: "HostBatches", first: int = 0, count: Optional[int] = None) -> int:
        """Pack the used prefixes of batches [first, first + count) on the device (tchgeo_pack_ragged) and copy them to
        the pinned buffers of `host` on the current stream: three D2H copies of exactly the used bytes (`rows` is not
        copied: it is arange(S, S + E), HostBatches.batch serves it from a cached host arange).  Asynchronous; returns
        the number of bytes that travel."""
        call = self._call
        return packed_to_host(self.samples, self.cols, self.edge_index, self.samples_len, self.edges_len,
                              int(call.cap_n[0]), int(call.cap_e[0]), call.device, host, first,
                              call.B - first if count is None else int(count), max_fanout=call.max_fanout)

    def relabeled(self, b):
        """-> (nodes, local) of batch b: nodes = seeds ++ every other id of samples at its first appearance,
        local[i] = index into nodes of samples[i] (negative_sampling.rs:20-47 semantic); the relabelled edges are
        (local[rows], local[cols])."""
        if self.nodes is None:
            raise ValueError("the sampler was created without relabel=True")
        return self.nodes[b, :int(self.nodes_len[b])], self.local[b, :int(self.samples_len[b])]


_host_arange_cache = {}


def host_arange(n: int) -> Tensor:
    """A cached host arange(n) (int64): `rows` of a homogeneous result is arange(S, S + E) for every batch
    (neighbor_sampling.rs:210-218: every sampled edge appends exactly one node), so the host serves it from a view of
    this instead of copying it back from the device."""
    t = _host_arange_cache.get("t")
    if t is None or t.numel() < n:
        t = torch.arange(max(int(n), 1 << 20), dtype=torch.int64)
        _host_arange_cache["t"] = t
    return t


_unpack_pool_cache = {}


def _unpack_pool():
    """ONE worker for every HostBatches of the process: groups are rebuilt one at a time, in arrival order, each with
    all of the caller's threads (several groups at once would only share the same cores and memory channels)."""
    import concurrent.futures
    if "pool" not in _unpack_pool_cache:
        _unpack_pool_cache["pool"] = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="tchgeo-unpack")
    return _unpack_pool_cache["pool"]


class HostBatches:
    """Host landing zone (+ its device staging buffers) for `SampledBatches.to_host`: the used prefixes of `samples`,
    `cols` and `edge_index` of a group of batches arrive packed back to back as int64 vectors.
    `fill` = fraction of the worst-case capacity to provide (sampled trees typically use about 0.65 of it).

    transport="plain": three D2H copies of the packed i64 vectors into pinned memory (24 B per edge on the bus).
    transport="compact": the group travels as i32 samples / edge positions and one u8 edge count per node (9 B per edge;
    tchgeo_pack_transport) into pinned staging buffers, and a worker thread rebuilds the i64 vectors in host memory with
    `threads` threads (tchgeo_host_unpack_transport) once the copies have arrived -- `wait()` (also called by `batch`)
    blocks until they are there.  Same bytes in the end; needs ids < 2^31 and fanouts <= 255.
    transport="hybrid": as compact, but `edge_index` travels as i64 straight into its pinned vector (13 B per edge on
    the bus, a third less for the host threads to write): the better balance when the host threads are the limit."""

    def __init__(self, num_batches: int, cap_samples: int, cap_edges: int, num_seeds: int, device, fill: float = 0.8,
                 transport: str = "plain", threads: Optional[int] = None):
        if transport not in ("plain", "compact", "hybrid"):
            raise ValueError("transport must be 'plain', 'compact' or 'hybrid'")
        self.B, self.S, self.device, self.transport = int(num_batches), int(num_seeds), device, transport
        self.cap_n = int(num_batches * cap_samples * fill) + 1
        self.cap_e = int(num_batches * cap_edges * fill) + 1
        host = lambda n, dt=torch.int64: torch.empty(n, dtype=dt).pin_memory()
        devb = lambda n, dt=torch.int64: torch.empty(n, dtype=dt, device=device)
        self._d_lens = devb(2 * self.B)
        self._d_off = devb(2 * (self.B + 1))
        # the lengths of a group reach the device by an asynchronous copy out of pinned memory: the source must stay
        # untouched until the copy has RUN, so consecutive transfers rotate through four small buffers, each guarded
        # by an event (one buffer would be overwritten by the next group while its copy still waits behind the
        # previous group's transfers)
        self._h_lens_ring = [host(2 * self.B) for _ in range(4)]
        self._h_lens_events = [None] * 4
        self._h_lens_next = 0
        self.n_off = self.e_off = None       # host offsets [count + 1] of the last to_host
        self.nbytes = 0
        self._pending = None
        if transport == "plain":
            self.samples, self.cols, self.edge_index = host(self.cap_n), host(self.cap_e), host(self.cap_e)
            self._d_samples, self._d_cols, self._d_eidx = devb(self.cap_n), devb(self.cap_e), devb(self.cap_e)
            return
        import os
        self.threads = int(threads) if threads else max(1, min(32, (os.cpu_count() or 2) - 1))
        self.samples = torch.empty(self.cap_n, dtype=torch.int64)
        self.cols = torch.empty(self.cap_e, dtype=torch.int64)
        self._h_s32, self._h_cnt = host(self.cap_n, torch.int32), host(self.cap_n, torch.uint8)
        self._d_s32, self._d_cnt = devb(self.cap_n, torch.int32), devb(self.cap_n, torch.uint8)
        if transport == "hybrid":
            self.edge_index, self._d_eidx = host(self.cap_e), devb(self.cap_e)
            self._h_e32 = self._d_e32 = None
        else:
            self.edge_index = torch.empty(self.cap_e, dtype=torch.int64)
            self._h_e32, self._d_e32 = host(self.cap_e, torch.int32), devb(self.cap_e, torch.int32)
        self._d_err, self._h_err = devb(1, torch.int32), host(1, torch.int32)
        self._event = torch.cuda.Event()
        self._pool = _unpack_pool()

    def wait(self):
        """compact transport: block until the worker has rebuilt the last group's vectors (no-op otherwise)"""
        fut, self._pending = self._pending, None
        if fut is not None:
            fut.result()

    def _unpack(self, n_off, e_off, count):
        self._event.synchronize()
        N.check(N.lib.tchgeo_status_from_error_word(int(self._h_err[0]) & 0xFFFFFFFF))
        hybrid = self._h_e32 is None
        N.check(N.lib.tchgeo_host_unpack_transport(_ptr(self._h_s32), None if hybrid else _ptr(self._h_e32), _ptr(self._h_cnt),
                                                   n_off.ctypes.data, e_off.ctypes.data, count, _ptr(self.samples),
                                                   _ptr(self.cols), None if hybrid else _ptr(self.edge_index), self.threads))

    def batch(self, i):
        """(samples, rows, cols, edge_index) of the i-th batch of the last transfer, as host views (plain transport:
        valid once the stream the transfer went to has been synchronised)"""
        self.wait()
        n0, n1, e0, e1 = int(self.n_off[i]), int(self.n_off[i + 1]), int(self.e_off[i]), int(self.e_off[i + 1])
        rows = host_arange(self.S + (e1 - e0))[self.S:self.S + (e1 - e0)]
        return self.samples[n0:n1], rows, self.cols[e0:e1], self.edge_index[e0:e1]


def packed_to_host(samples, cols, edge_index, samples_len, edges_len, cap_n, cap_e, device, host: "HostBatches",
                   first: int, count: int, max_fanout: Optional[int] = None) -> int:
    """Shared by SampledBatches.to_host and the partitioned plans' results (same padded [B, capacity] layout)."""
    if count > host.B:
        raise ValueError("HostBatches is smaller than the group of batches")
    compact, hybrid = host.transport in ("compact", "hybrid"), host.transport == "hybrid"
    if compact and max_fanout is not None and max_fanout > 255:
        raise ValueError("transport='compact' carries one u8 edge count per node: fanouts above 255 need transport='plain'")
    host.wait()                              # the staging buffers of the previous group are free again
    ns, ne = samples_len[first:first + count], edges_len[first:first + count]
    n_off = np.ascontiguousarray(np.concatenate([[0], np.cumsum(ns)]), dtype=np.int64)
    e_off = np.ascontiguousarray(np.concatenate([[0], np.cumsum(ne)]), dtype=np.int64)
    if n_off[-1] > host.cap_n or e_off[-1] > host.cap_e:
        raise MemoryError("HostBatches too small for this group: create it with a larger `fill`")
    slot = host._h_lens_next
    host._h_lens_next = (slot + 1) % len(host._h_lens_ring)
    if host._h_lens_events[slot] is not None:
        host._h_lens_events[slot].synchronize()      # four transfers ago: long done
    h = host._h_lens_ring[slot]
    h[:count].copy_(torch.from_numpy(np.ascontiguousarray(ns)))
    h[host.B:host.B + count].copy_(torch.from_numpy(np.ascontiguousarray(ne)))
    nt, et = int(n_off[-1]), int(e_off[-1])
    with torch.cuda.device(device):
        stream = _stream(device)
        host._d_lens.copy_(h, non_blocking=True)
        if host._h_lens_events[slot] is None:
            host._h_lens_events[slot] = torch.cuda.Event()
        host._h_lens_events[slot].record(torch.cuda.current_stream(device))
        if compact:
            host._d_err.zero_()
            N.check(N.lib.tchgeo_pack_transport(_ptr(samples[first]), samples.shape[1], _ptr(cols[first]),
                                                _ptr(edge_index[first]), cols.shape[1], _ptr(host._d_lens),
                                                _ptr(host._d_lens[host.B:]), count, int(cap_n), int(cap_e), _ptr(host._d_s32),
                                                None if hybrid else _ptr(host._d_e32), _ptr(host._d_cnt), nt, _ptr(host._d_off),
                                                _ptr(host._d_off[host.B + 1:]), _ptr(host._d_err), stream))
            host._h_s32[:nt].copy_(host._d_s32[:nt], non_blocking=True)
            host._h_cnt[:nt].copy_(host._d_cnt[:nt], non_blocking=True)
            if hybrid:
                N.check(N.lib.tchgeo_pack_ragged(_ptr(edge_index[first]), edge_index.shape[1], _ptr(host._d_lens[host.B:]), 1,
                                                 count, int(cap_e), _ptr(host._d_eidx), _ptr(host._d_off[host.B + 1:]), stream))
                host.edge_index[:et].copy_(host._d_eidx[:et], non_blocking=True)
            else:
                host._h_e32[:et].copy_(host._d_e32[:et], non_blocking=True)
            host._h_err.copy_(host._d_err, non_blocking=True)
            host._event.record(torch.cuda.current_stream(device))
            host._pending = host._pool.submit(host._unpack, n_off, e_off, count)
            host.nbytes = 5 * nt + (8 if hybrid else 4) * et + 4
        else:
            for src, lens_at, dst, off_at, cap in ((samples, 0, host._d_samples, 0, cap_n),
                                                   (cols, host.B, host._d_cols, host.B + 1, cap_e),
                                                   (edge_index, host.B, host._d_eidx, host.B + 1, cap_e)):
                N.check(N.lib.tchgeo_pack_ragged(_ptr(src[first]), src.shape[1], _ptr(host._d_lens[lens_at:]), 1, count,
                                                 int(cap), _ptr(dst), _ptr(host._d_off[off_at:]), stream))
            host.samples[:nt].copy_(host._d_samples[:nt], non_blocking=True)
            host.cols[:et].copy_(host._d_cols[:et], non_blocking=True)
            host.edge_index[:et].copy_(host._d_eidx[:et], non_blocking=True)
            host.nbytes = 8 * (nt + 2 * et)
    host.n_off, host.e_off = n_off, e_off
    return host.nbytes


class HomogenousSampler:
    """Reusable plan for repeated batched sampling over one graph (extension; not in the reference).
    Output buffers and workspace are allocated once; `sample(inputs)` enqueues B batches in H launches."""

    def __init__(self, col_ptrs, row_indices, num_batches, seeds_per_batch, num_neighbors, sampler=None, relabel=False,
                 filter=None):
        """relabel=True adds the dedup + insertion-order relabel stage to every step (results: SampledBatches.nodes /
        .local / .nodes_len, `relabeled(b)`); the reference-layout outputs are unchanged.
        filter: a TemporalEdgeFilter (python.rs:137-168); every `sample` call then takes `inputs_state` [B, S]."""
        dev = col_ptrs.device
        self._inputs = torch.zeros((num_batches, seeds_per_batch), dtype=torch.int64, device=dev)
        self._state = torch.zeros_like(self._inputs) if filter is not None else None
        self._call = _homogenous_call(col_ptrs, row_indices, self._inputs, num_neighbors, sampler,
                                      (filter, self._state) if filter is not None else None, batched=True, relabel=relabel)

    def sample(self, inputs: Tensor, seed: Optional[int] = None, batch_base: int = 0,
               timed: bool = False, inputs_state: Optional[Tensor] = None) -> SampledBatches:
        """inputs: [B, S] i64 on the device or in (pinned) host memory.  With timed=True the result
        carries `launch_ms`, the device duration of each hop kernel (CUDA events on the current stream)."""
        if not isinstance(inputs, Tensor) or inputs.dtype != torch.int64:
            raise ValueError("inputs must be an int64 tensor")
        if tuple(inputs.shape) != tuple(self._inputs.shape):
            raise ValueError(f"inputs must have shape {tuple(self._inputs.shape)}")
        self._inputs.copy_(inputs, non_blocking=True)
        if self._state is not None:
            if inputs_state is None or tuple(inputs_state.shape) != tuple(self._state.shape):
                raise N.ReferencePanic("inputs_state must have one entry per input")  # states[i] index panic (quirk Q10)
            self._state.copy_(inputs_state, non_blocking=True)
        self._call.run(seed=_rng_get() if seed is None else seed, batch_base=batch_base, timed=timed)
        res = SampledBatches(self._call)
        res.launch_ms = getattr(self._call, "launch_ms", None) if timed else None
        return res

    def sample_async(self, inputs: Tensor, seed: Optional[int] = None, batch_base: int = 0) -> "HomogenousSampler":
        """Enqueue `sample(inputs)` on the current stream without waiting for it; `result()` waits for that stream
        and returns the SampledBatches.  With two plans on two streams (enqueue plan B, then take plan A's
        result) the device never idles between steps: the length read-back of one step overlaps the kernels
        of the next.  The plan's buffers belong to the pending step until `result()` has returned."""
        if not isinstance(inputs, Tensor) or inputs.dtype != torch.int64:
            raise ValueError("inputs must be an int64 tensor")
        if tuple(inputs.shape) != tuple(self._inputs.shape):
            raise ValueError(f"inputs must have shape {tuple(self._inputs.shape)}")
        self._inputs.copy_(inputs, non_blocking=True)
        self._call.run(seed=_rng_get() if seed is None else seed, batch_base=batch_base, collect=False)
        return self

    def result(self) -> SampledBatches:
        self._call.collect()
        res = SampledBatches(self._call)
        res.launch_ms = None
        return res

    @property
    def num_launches(self):
        """kernel launches of one step (hop kernels, the seed-length fill, the relabel stage's kernels)"""
        return self._call.num_launches


def neighbor_sampling_homogenous_batched(col_ptrs, row_indices, inputs, num_neighbors, sampler=None,
                                         seed: Optional[int] = None, batch_base: int = 0,
                                         relabel: bool = False) -> SampledBatches:
    """Extension: inputs [B, S] -> B independent neighbor_sampling_homogenous results in one call."""
    call = _homogenous_call(col_ptrs, row_indices, inputs, num_neighbors, sampler, None, batched=True, relabel=relabel)
    call.run(seed=_rng_get() if seed is None else seed, batch_base=batch_base)
    return SampledBatches(call)


def rel_key(edge_type) -> str:
    """neighbor_sampling.rs:255-258"""
    return "{}__{}__{}".format(*edge_type)


def _heterogenous_call(node_types, edge_types, col_ptrs, row_indices, inputs, num_neighbors, num_hops, sampler,
                       filter, num_batches=None, relabel=False):
    """Argument checking + plan for the heterogeneous sampler.  inputs[t]: [S_t] (single call) or
    [B, S_t] when num_batches is given."""
    flt = _extract_filter(filter, hetero=True)
    node_types = list(node_types)
    edge_types = [tuple(e) for e in edge_types]
    tix = {t: i for i, t in enumerate(node_types)}
    rels = [rel_key(e) for e in edge_types]
    # graphs are built from col_ptrs.keys() (python.rs:293-300): every key needs row_indices and an edge type
    for r in col_ptrs:
        if r not in row_indices:
            raise KeyError(r)
    for r in num_neighbors:
        if r not in rels:
            raise KeyError(r)  # to_edge_types[rel_type] panics, neighbor_sampling.rs:296
        if r not in col_ptrs:
            raise KeyError(r)  # graphs[rel_type] panics, :310
    dev = None
    for t in list(col_ptrs.values()):
        dev = t.device
        break
    if dev is None:
        raise ValueError("col_ptrs is empty")
    kind, wdict = _extract_sampler(sampler, hetero=True)
    cp, ri, ws = [], [], []
    for r in rels:
        if r in col_ptrs:
            cp.append(_check(col_ptrs[r], torch.int64, f"col_ptrs[{r}]", dev))
            ri.append(_check(row_indices[r], torch.int64, f"row_indices[{r}]", dev))
        else:
            cp.append(None)
            ri.append(None)
        if kind == N.SAMPLER_WEIGHTED:
            ws.append(_check(wdict[r], torch.float64, f"weights[{r}]", dev) if r in wdict else None)
    active = [1 if r in num_neighbors else 0 for r in rels]
    fan = np.zeros((len(rels), max(num_hops, 1)), dtype=np.int64)
    for i, r in enumerate(rels):
        if r in num_neighbors:
            ks = [int(k) for k in num_neighbors[r]]
            if len(ks) < num_hops:
                raise IndexError("num_neighbors[%s] has fewer than num_hops entries" % r)  # :295 index panic
            fan[i, :num_hops] = ks[:num_hops]
    B = 1 if num_batches is None else int(num_batches)
    inp, seeds = [], []
    for t in node_types:
        if t in inputs:
            x = _as_seed_matrix(inputs[t], dev, f"inputs[{t}]").reshape(B, -1)
            inp.append(x)
            seeds.append(x.shape[1])
        else:
            inp.append(None)
            seeds.append(0)
    filt = None
    if flt is not None:
        mode, forward, window, tsd, std = flt
        ts = [(_check(tsd[r], torch.int64, f"timestamps[{r}]", dev) if r in tsd else None) for r in rels]
        st = []
        for t, x in zip(node_types, inp):
            if x is None:
                st.append(None)
                continue
            if t not in std:
                raise N.ReferencePanic(f"inputs_state[{t}] is missing")  # states[node_type] stays empty -> index panic
            y = _as_seed_matrix(std[t], dev, f"inputs_state[{t}]").reshape(B, -1)
            if y.shape != x.shape:
                raise N.ReferencePanic(f"inputs_state[{t}] must have one entry per input")
            st.append(y)
        filt = (mode, forward, window, ts, st)
    call = _Call(dev, [tix[e[0]] for e in edge_types], [tix[e[2]] for e in edge_types], cp, ri,
                 ws if kind == N.SAMPLER_WEIGHTED else None, fan[:, :num_hops].reshape(-1) if num_hops > 0 else [],
                 active, inp, seeds, B, num_hops, kind, 0, filt=filt, relabel=relabel)
    call.meta = (node_types, rels, [r in col_ptrs for r in rels], active, num_hops)
    call.seed_inputs = inp
    return call


def _hetero_batch(call, b):
    node_types, rels, present, active, num_hops = call.meta
    out_s = {t: call.samples[i][b, :int(call.samples_len[b, i])] for i, t in enumerate(node_types)}
    out_r, out_c, out_e, out_lo = {}, {}, {}, {}
    for i, r in enumerate(rels):
        if not present[i]:
            continue  # outputs are keyed by graphs.keys(), neighbor_sampling.rs:280-285
        ne = int(call.edges_len[b, i])
        out_r[r], out_c[r], out_e[r] = call.rows[i][b, :ne], call.cols[i][b, :ne], call.eidx[i][b, :ne]
        out_lo[r] = ([tuple(int(x) for x in call.layer_offsets[b, i, h]) for h in range(num_hops)]
                     if active[i] else [])
    return out_s, out_r, out_c, out_e, out_lo


def neighbor_sampling_heterogenous(
        node_types: List[str],
        edge_types: List[Tuple[str, str, str]],
        col_ptrs: Dict[str, Tensor],
        row_indices: Dict[str, Tensor],
        inputs: Dict[str, Tensor],
        num_neighbors: Dict[str, List[int]],
        num_hops: int,
        sampler=None,
        filter=None,
):
    """python.rs:273-395.  Returns (samples{node_type}, rows{rel}, cols{rel}, edge_index{rel},
    layer_offsets{rel}).  Relations are visited in `edge_types` order (the reference iterates a
    HashMap, quirk Q6)."""
    call = _heterogenous_call(node_types, edge_types, col_ptrs, row_indices, inputs, num_neighbors, num_hops, sampler,
                              filter)
    call.run(seed=_rng_get())
    return _hetero_batch(call, 0)


class HeterogenousSampler:
    """Reusable plan for batched heterogeneous sampling (extension): `sample({type: [B, S_t]})` runs B
    independent neighbor_sampling_heterogenous calls with one launch per (hop, relation)."""

    def __init__(self, node_types, edge_types, col_ptrs, row_indices, num_batches, seeds_per_batch: Dict[str, int],
                 num_neighbors, num_hops, sampler=None, relabel=False):
        dev = next(iter(col_ptrs.values())).device
        proto = {t: torch.zeros((num_batches, int(s)), dtype=torch.int64, device=dev) for t, s in seeds_per_batch.items()}
        self._call = _heterogenous_call(node_types, edge_types, col_ptrs, row_indices, proto, num_neighbors, num_hops,
                                        sampler, None, num_batches=num_batches, relabel=relabel)
        self._proto = proto
        self.num_batches = num_batches

    def sample(self, inputs: Dict[str, Tensor], seed: Optional[int] = None, batch_base: int = 0, timed: bool = False):
        for t, buf in self._proto.items():
            if tuple(inputs[t].shape) != tuple(buf.shape):
                raise ValueError(f"inputs[{t}] must have shape {tuple(buf.shape)}")
            buf.copy_(inputs[t], non_blocking=True)
        self._call.run(seed=_rng_get() if seed is None else seed, batch_base=batch_base, timed=timed)
        return self

    def sample_async(self, inputs: Dict[str, Tensor], seed: Optional[int] = None, batch_base: int = 0):
        """Enqueue `sample(inputs)` on the current stream without waiting; `result()` waits for that step only.  Two plans
        on two streams keep the device busy while the host reads the previous step's lengths."""
        for t, buf in self._proto.items():
            if tuple(inputs[t].shape) != tuple(buf.shape):
                raise ValueError(f"inputs[{t}] must have shape {tuple(buf.shape)}")
            buf.copy_(inputs[t], non_blocking=True)
        self._call.run(seed=_rng_get() if seed is None else seed, batch_base=batch_base, collect=False)
        return self

    def result(self):
        self._call.collect()
        return self

    @property
    def launch_ms(self):
        return getattr(self._call, "launch_ms", None)

    @property
    def edges_len(self):
        return self._call.edges_len  # [B, R]

    @property
    def samples_len(self):
        return self._call.samples_len  # [B, T]

    @property
    def layer_offsets(self):
        return self._call.layer_offsets  # [B, R, H, 3]; -1 where a relation is not sampled

    def batch(self, b):
        return _hetero_batch(self._call, b)

    def relabeled(self, b):
        """-> {node_type: (nodes, local)} of batch b (K7 per node type; needs relabel=True)"""
        c = self._call
        if c.nodes is None:
            raise ValueError("the sampler was created without relabel=True")
        return {t: (c.nodes[i][b, :int(c.nodes_len[b, i])], c.local[i][b, :int(c.samples_len[b, i])])
                for i, t in enumerate(c.meta[0])}

    @property
    def num_launches(self):
        return self._call.num_launches


# ---------------------------------------------------------------------------------------------
# random_walk (python.rs:583-608 -> random_walk.rs:10-75)
# ---------------------------------------------------------------------------------------------
def random_walk(row_ptrs: Tensor, col_indices: Tensor, start: Tensor, walk_length: int, p: float, q: float,
                *, seed: Optional[int] = None, walker_base: int = 0, return_attempts: bool = False):
    _check(row_ptrs, torch.int64, "row_ptrs")
    dev = row_ptrs.device
    _check(col_indices, torch.int64, "col_indices", dev)
    start = _as_seed_matrix(start, dev, "start").reshape(-1)
    S = start.numel()
    walks = torch.empty((S, walk_length + 1), dtype=torch.int64, device=dev)
    stats = torch.empty(2, dtype=torch.int64, device=dev)
    attempts = ctypes.c_int64(0)
    graph = _graph_cache.get([row_ptrs], [col_indices])   # owns the int32 replica of col_indices
    with torch.cuda.device(dev):
        st = N.lib.tchgeo_random_walk_graph(graph.handle, 0, _ptr(start), S, int(walk_length), float(p), float(q),
                                            _rng_get() if seed is None else seed, int(walker_base), _ptr(walks),
                                            _ptr(stats), ctypes.addressof(attempts), _stream(dev))
    N.check(st)
    return (walks, attempts.value) if return_attempts else walks


# ---------------------------------------------------------------------------------------------
# negative neighbour sampling (python.rs:689-783 -> negative_sampling.rs:6-131)
# ---------------------------------------------------------------------------------------------
def _negative(node_types, edge_types, row_ptrs, col_indices, sizes, inputs, num_neg, try_count, inbound, seed):
    node_types = list(node_types)
    edge_types = [tuple(e) for e in edge_types]
    tix = {t: i for i, t in enumerate(node_types)}
    rels = [rel_key(e) for e in edge_types]
    T, R = len(node_types), len(rels)
    for r in row_ptrs:  # graphs are built from row_ptrs.keys(), python.rs:748-753
        if r not in col_indices or r not in sizes:
            raise KeyError(r)
    for t in inputs:
        if t not in tix:
            raise KeyError(t)
    dev = next(iter(row_ptrs.values())).device
    keep = []

    def host(arr, dtype):
        x = np.ascontiguousarray(np.asarray(arr, dtype=dtype))
        keep.append(x)
        return x

    def ptr_table(tensors):
        x = np.array([(t.data_ptr() if t is not None else 0) for t in tensors], dtype=np.uint64)
        keep.append(x)
        keep.append(list(tensors))
        return x

    rp, ci, nrows, ncount = [], [], [], []
    for r in rels:
        if r not in row_ptrs:
            raise KeyError(r)  # &graphs[rel_type] panics when the relation is drawn, negative_sampling.rs:105
        rp.append(_check(row_ptrs[r], torch.int64, f"row_ptrs[{r}]", dev))
        ci.append(_check(col_indices[r], torch.int64, f"col_indices[{r}]", dev))
        nrows.append(rp[-1].numel() - 1)
        ncount.append(_size_tuple(sizes[r])[1])
    inp = [(_as_seed_matrix(inputs[t], dev, f"inputs[{t}]").reshape(-1) if t in inputs else None) for t in node_types]
    a = N.NegativeArgs()
    a.num_node_types, a.num_rels = T, R
    a.rel_src = host([tix[e[0]] for e in edge_types], np.int32).ctypes.data
    a.rel_dst = host([tix[e[2]] for e in edge_types], np.int32).ctypes.data
    a.row_ptrs, a.col_indices = ptr_table(rp).ctypes.data, ptr_table(ci).ctypes.data
    a.num_rows, a.node_count = host(nrows, np.int64).ctypes.data, host(ncount, np.int64).ctypes.data
    a.inputs = ptr_table(inp).ctypes.data
    a.num_inputs = host([(x.numel() if x is not None else 0) for x in inp], np.int64).ctypes.data
    a.num_neg, a.try_count, a.inbound, a.seed = int(num_neg), int(try_count), 1 if inbound else 0, seed
    if a.num_neg < 0 or a.try_count < 0:
        num_neg, try_count = max(a.num_neg, 0), max(a.try_count, 0)  # `for _ in 0..n` with n < 0 is an empty loop
        a.num_neg, a.try_count = num_neg, try_count
    cap_n, cap_e = np.zeros(T, dtype=np.int64), np.zeros(R, dtype=np.int64)
    N.check(N.lib.tchgeo_negative_sampling_capacity(ctypes.byref(a), cap_n.ctypes.data, cap_e.ctypes.data))
    i64 = dict(dtype=torch.int64, device=dev)
    samples = [torch.empty(int(c), **i64) for c in cap_n]
    rows = [torch.empty(int(c), **i64) for c in cap_e]
    cols = [torch.empty(int(c), **i64) for c in cap_e]
    a.samples, a.rows, a.cols = ptr_table(samples).ctypes.data, ptr_table(rows).ctypes.data, ptr_table(cols).ctypes.data
    slen, elen = np.zeros(T, dtype=np.int64), np.zeros(R, dtype=np.int64)
    a.samples_len, a.edges_len = slen.ctypes.data, elen.ctypes.data
    ws_bytes = N.lib.tchgeo_negative_sampling_workspace_bytes(ctypes.byref(a))
    ws = torch.empty(max(int(ws_bytes), 1), dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
    with torch.cuda.device(dev):
        a.stream = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib.tchgeo_negative_sampling(ctypes.byref(a)))
    out_s = {t: samples[i][:int(slen[i])] for i, t in enumerate(node_types)}
    out_r = {r: rows[i][:int(elen[i])] for i, r in enumerate(rels)}
    out_c = {r: cols[i][:int(elen[i])] for i, r in enumerate(rels)}
    counts = {t: (inp[i].numel() if inp[i] is not None else 0) for i, t in enumerate(node_types)}
    return out_s, out_r, out_c, counts


def negative_sample_neighbors_homogenous(row_ptrs: Tensor, col_indices: Tensor, graph_size: Tuple[int, int],
                                         inputs: Tensor, num_neg: int, try_count: int, *,
                                         seed: Optional[int] = None) -> Tuple[Tensor, Tensor, Tensor, int]:
    """python.rs:689-721.  Returns (samples, rows, cols, sample_count): samples = inputs ++ accepted negatives at first
    appearance, (rows[e], cols[e]) index `samples`, sample_count = len(inputs)."""
    s, r, c, n = _negative(["n"], [("n", "e", "n")], {"n__e__n": row_ptrs}, {"n__e__n": col_indices},
                           {"n__e__n": graph_size}, {"n": inputs}, num_neg, try_count, False,
                           _rng_get() if seed is None else seed)
    return s["n"], r["n__e__n"], c["n__e__n"], n["n"]


def negative_sample_neighbors_heterogenous(node_types: List[str], edge_types: List[Tuple[str, str, str]],
                                           row_ptrs: Dict[str, Tensor], col_indices: Dict[str, Tensor],
                                           sizes: Dict[str, Tuple[int, int]], inputs: Dict[str, Tensor], num_neg: int,
                                           try_count: int, inbound: bool, *, seed: Optional[int] = None):
    """python.rs:723-783.  Returns (samples{node_type}, rows{rel}, cols{rel}, sample_count{node_type}).  Node types
    are visited in `node_types` order and relations in `edge_types` order (the reference iterates HashMaps)."""
    return _negative(node_types, edge_types, row_ptrs, col_indices, sizes, inputs, num_neg, try_count, inbound,
                     _rng_get() if seed is None else seed)


def tempo_random_walk(row_ptrs: Tensor, col_indices: Tensor, node_timestamps: Tensor, edge_timestamps: Tensor,
                      start: Tensor, start_timestamps: Tensor, walk_length: int, window: Tuple[int, int], *,
                      seed: Optional[int] = None, walker_base: int = 0) -> Tuple[Tensor, Tensor]:
    """python.rs:610-643 -> random_walk.rs:80-158.  Returns (walks, walk_timestamps), both [S, walk_length]."""
    _check(row_ptrs, torch.int64, "row_ptrs")
    dev = row_ptrs.device
    _check(col_indices, torch.int64, "col_indices", dev)
    _check(node_timestamps, torch.int64, "node_timestamps", dev)
    _check(edge_timestamps, torch.int64, "edge_timestamps", dev)
    if edge_timestamps.numel() < col_indices.numel():
        raise N.ReferencePanic("edge_timestamps is shorter than col_indices")  # get_range slices out of bounds
    start = _as_seed_matrix(start, dev, "start").reshape(-1)
    start_timestamps = _as_seed_matrix(start_timestamps, dev, "start_timestamps").reshape(-1)
    S = start.numel()
    if start_timestamps.numel() < S:
        raise N.ReferencePanic("start_timestamps is shorter than start")
    L = int(walk_length)
    if L < 0:
        raise RuntimeError("walk_length must not be negative")  # Tensor::full with a negative size fails in libtorch
    walks = torch.empty((S, L), dtype=torch.int64, device=dev)
    walks_ts = torch.empty((S, L), dtype=torch.int64, device=dev)
    scratch = torch.empty(1, dtype=torch.int32, device=dev)
    w0, w1 = int(window[0]), int(window[1])
    with torch.cuda.device(dev):
        N.check(N.lib.tchgeo_tempo_random_walk(_ptr(row_ptrs), row_ptrs.numel() - 1, _ptr(col_indices), _ptr(node_timestamps),
                                               node_timestamps.numel(), _ptr(edge_timestamps), _ptr(start),
                                               _ptr(start_timestamps), S, L, w0, w1, _rng_get() if seed is None else seed,
                                               int(walker_base), _ptr(walks), _ptr(walks_ts), _ptr(scratch), _stream(dev)))
    return walks, walks_ts


# ---------------------------------------------------------------------------------------------
# downstream gather (the step after the sampler in every loader: x[samples], edge_attr[perm[edge_index]])
# ---------------------------------------------------------------------------------------------
def gather_rows(src: Tensor, index: Tensor, out: Optional[Tensor] = None, validate: bool = True) -> Tensor:
    """src[index] along dim 0 for a contiguous CUDA tensor of any dtype and an int64 index vector
    (examples/neighbor_sampling.py:21-24 / PyG filter_data).  Out-of-range indices raise (no negative wrap-around).
    out: optional preallocated result; validate=False skips the error read-back, which makes the call asynchronous
    (use it when the index comes straight from the sampler, whose ids are in range by construction)."""
    if not isinstance(src, Tensor) or not src.is_cuda:
        raise ValueError("src must be a CUDA tensor")
    if not src.is_contiguous() or src.dim() < 1:
        raise ValueError("src must be contiguous with at least one dimension")
    dev = src.device
    _check(index, torch.int64, "index", dev)
    if index.dim() != 1:
        raise ValueError("index must be a vector")
    n = index.numel()
    shape = (n,) + tuple(src.shape[1:])
    if out is None:
        out = torch.empty(shape, dtype=src.dtype, device=dev)
    elif tuple(out.shape) != shape or out.dtype != src.dtype or out.device != dev or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous {src.dtype} tensor of shape {shape} on {dev}")
    row_bytes = (src.numel() // src.shape[0]) * src.element_size() if src.shape[0] > 0 else 0
    if src.shape[0] == 0 and n > 0:
        raise N.ReferencePanic("index out of range (src has no rows)")
    scratch = torch.empty(1, dtype=torch.int32, device=dev) if validate else None
    with torch.cuda.device(dev):
        N.check(N.lib.tchgeo_gather_rows(_ptr(src), src.shape[0], row_bytes, _ptr(index), n, _ptr(out), _ptr(scratch),
                                         _stream(dev)))
    return out


# ---------------------------------------------------------------------------------------------
# dedup + relabel stage (additive; semantic of negative_sampling.rs:20-47)
# ---------------------------------------------------------------------------------------------
def unique_relabel(samples: Tensor, num_seeds: int) -> Tuple[Tensor, Tensor]:
    """-> (nodes, local): nodes = seeds ++ other ids in first-appearance order; local[i] indexes nodes."""
    _check(samples, torch.int64, "samples")
    dev = samples.device
    n = samples.numel()
    nodes = torch.empty(n, dtype=torch.int64, device=dev)
    local = torch.empty(n, dtype=torch.int64, device=dev)
    ws_bytes = N.lib.tchgeo_unique_relabel_workspace_bytes(n)
    ws = torch.empty(max(int(ws_bytes), 1), dtype=torch.uint8, device=dev)
    num = ctypes.c_int64(0)
    with torch.cuda.device(dev):
        st = N.lib.tchgeo_unique_relabel(_ptr(samples), n, int(num_seeds), _ptr(nodes), _ptr(local),
                                         ctypes.addressof(num), _ptr(ws), ws_bytes, _stream(dev))
    N.check(st)
    return nodes[:num.value], local


def unique_relabel_batched(samples: Tensor, lens: Tensor, num_seeds: int, key32: bool = False,
                           id_bound: Optional[int] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """B trees in padded rows: samples [B, stride], lens [B] (device) -> (nodes [B, stride], local [B, stride],
    nodes_len [B]).  The stage HomogenousSampler(relabel=True) runs after the hops, exposed on its own.
    id_bound: every id is in [0, id_bound) (the node count; lets the stage use direct-address tables); key32=True is
    id_bound = 2**32 - 1; neither: any non-negative i64."""
    bound = int(id_bound) if id_bound is not None else (0xFFFFFFFF if key32 else 0)
    if not 0 <= bound <= 0xFFFFFFFF:
        raise ValueError("id_bound must be in [0, 2**32 - 1]")
    _check(samples, torch.int64, "samples")
    dev = samples.device
    _check(lens, torch.int64, "lens", dev)
    if samples.dim() != 2 or lens.numel() != samples.shape[0]:
        raise ValueError("samples must be [B, stride] and lens [B]")
    B, stride = samples.shape
    nodes, local = torch.empty_like(samples), torch.empty_like(samples)
    nodes_len = torch.zeros(B, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = N.lib.tchgeo_unique_relabel_batched_workspace_bytes(B, stride, bound)
    ws = torch.empty(max(int(ws_bytes), 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.tchgeo_unique_relabel_batched(_ptr(samples), stride, _ptr(lens), B, int(num_seeds), stride,
                                                    bound, _ptr(nodes), _ptr(local), _ptr(nodes_len), _ptr(ws),
                                                    ws_bytes, _ptr(err), _stream(dev)))
    N.check(N.lib.tchgeo_status_from_error_word(int(err.item()) & 0xFFFFFFFF))
    return nodes, local, nodes_len
