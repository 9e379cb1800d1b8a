"""CPU-only: host-side logic of the Python mirror (argument extraction, error behaviour, RNG forking,
capacity planning, sharding).  Mirrors src/python.rs:107-168 and src/utils/tensor.rs:10-70."""
import ctypes

import numpy as np
import pytest
import torch

import tch_geometric as thg
from tch_geometric import _native as N
from tch_geometric import ops
from tch_geometric.sharding import shard_range
from oracle import oracle as O


def test_reference_surface_is_present():
    for name in ("to_csc", "to_csr", "neighbor_sampling_homogenous", "neighbor_sampling_heterogenous", "random_walk",
                 "UniformEdgeSampler", "WeightedEdgeSampler", "TemporalEdgeFilter", "TEMPORAL_SAMPLE_STATIC"):
        assert hasattr(thg, name)
    assert thg.UniformEdgeSampler().with_replacement is False
    assert (thg.TEMPORAL_SAMPLE_STATIC, thg.TEMPORAL_SAMPLE_RELATIVE, thg.TEMPORAL_SAMPLE_DYNAMIC) == (0, 1, 2)


def test_cpu_tensors_are_rejected_not_silently_computed():
    """No CPU fallback: the reference's InvalidDevice error becomes 'must be on Cuda device'."""
    p = torch.zeros(3, dtype=torch.int64)
    with pytest.raises(ValueError, match="Cuda"):
        thg.neighbor_sampling_homogenous(p, p, p, [2])
    with pytest.raises(ValueError, match="Cuda"):
        thg.to_csc(torch.zeros((2, 3), dtype=torch.int64), 3)
    with pytest.raises(ValueError, match="Cuda"):
        thg.random_walk(p, p, p, 3, 1.0, 1.0)
    with pytest.raises(ValueError, match="Cuda"):
        thg.unique_relabel(p, 1)


def test_sampler_extraction_order():
    """derive(FromPyObject) tries Uniform{with_replacement: bool} first, then Weighted{weights}."""
    assert ops._extract_sampler(None, False) == (N.SAMPLER_UNIFORM, None)
    assert ops._extract_sampler(thg.UniformEdgeSampler(True), False)[0] == N.SAMPLER_UNIFORM_REPLACE
    assert ops._extract_sampler(thg.UniformEdgeSampler(False), True)[0] == N.SAMPLER_UNIFORM
    w = torch.ones(3, dtype=torch.float64)
    assert ops._extract_sampler(thg.WeightedEdgeSampler(w), False) == (N.SAMPLER_WEIGHTED, w)
    with pytest.raises(ValueError, match="homogenous"):
        ops._extract_sampler(thg.WeightedEdgeSampler({"a": w}), False)
    with pytest.raises(ValueError, match="heterogenous"):
        ops._extract_sampler(thg.WeightedEdgeSampler(w), True)
    with pytest.raises(TypeError):
        ops._extract_sampler(object(), False)

    class Both:  # an object with both attributes is a uniform sampler, as in the reference
        with_replacement = True
        weights = w
    assert ops._extract_sampler(Both(), False)[0] == N.SAMPLER_UNIFORM_REPLACE


def test_graph_size_extraction():
    assert ops._size_tuple(5) == (5, 5) and ops._size_tuple((3, 4)) == (3, 4)
    with pytest.raises(ValueError):
        ops._size_tuple((1, 2, 3))


def test_rng_forks_a_child_per_call_and_reseeds():
    thg.rng_reseed(42)
    a = [ops._rng_get() for _ in range(3)]
    thg.rng_reseed(42)
    b = [ops._rng_get() for _ in range(3)]
    assert a == b and len(set(a)) == 3
    assert a[0] == ops.splitmix64(42)[1]
    # splitmix64 known answer (seed 0 -> first output), the same expansion rand uses for seed_from_u64
    assert ops.splitmix64(0)[1] == 0xE220A8397B1DCDAF


def _capacity(T, R, H, rel_src, rel_dst, fan, seeds, active=None, kind=0, B=1):
    a = N.SamplingArgs()
    keep = [np.asarray(rel_src, dtype=np.int32), np.asarray(rel_dst, dtype=np.int32), np.asarray(fan, dtype=np.int64),
            np.asarray(seeds, dtype=np.int64), np.asarray(active if active is not None else [1] * R, dtype=np.uint8)]
    a.num_node_types, a.num_rels, a.num_hops, a.sampler_kind, a.num_batches = T, R, H, kind, B
    a.rel_src, a.rel_dst, a.fanouts, a.seeds_per_batch, a.rel_active = (k.ctypes.data for k in keep)
    cn, ce = np.zeros(T, dtype=np.int64), np.zeros(R, dtype=np.int64)
    N.check(N.lib.tchgeo_neighbor_sampling_capacity(ctypes.byref(a), cn.ctypes.data, ce.ctypes.data))
    return cn.tolist(), ce.tolist()


def test_capacity_planner_matches_reference_recurrence():
    assert _capacity(1, 1, 2, [0], [0], [5, 5], [34]) == ([34 + 170 + 850], [170 + 850])
    assert _capacity(1, 1, 0, [0], [0], [], [7]) == ([7], [0])
    # hetero: types a,b ; rels r0: a->b (src a, dst b), r1: b->b ; seeds only on b
    cn, ce = _capacity(2, 2, 2, [0, 1], [1, 1], [2, 3, 4, 5], [0, 10])
    # hop0: r0 adds 20 to a, r1 adds 40 to b ; hop1: frontier a=20 (no relation has dst a), b=40: r0 adds 120 to a, r1 adds 200 to b
    assert (cn, ce) == ([140, 250], [140, 240])
    # inactive relation contributes nothing
    cn, ce = _capacity(2, 2, 2, [0, 1], [1, 1], [2, 3, 4, 5], [0, 10], active=[1, 0])
    assert (cn, ce) == ([20, 10], [20, 0])


def test_capacity_bounds_the_oracle(karate):
    ei, n = karate
    ptrs, idx, _ = O.to_csc(ei, n)
    for fan in ([5, 5], [17, 17], [3, 2, 4]):
        for sampler in (None, ("uniform", True)):
            s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, np.arange(34), fan, sampler=sampler)
            cn, ce = _capacity(1, 1, len(fan), [0], [0], fan, [34], kind=1 if sampler else 0)
            assert len(s) <= cn[0] and len(r) <= ce[0]


def test_filter_extraction():
    """python.rs:137-168 + the dispatch at :219-249 (unknown modes fall through to IdentityFilter)."""
    p = torch.zeros(3, dtype=torch.int64)
    assert ops._extract_filter(None, False) is None
    f = ops._extract_filter((thg.TemporalEdgeFilter((0, 2), p), p), False)
    assert f[0] == 1 and f[1] is False and f[2] == (0, 2)
    f = ops._extract_filter((thg.TemporalEdgeFilter((-1, 5), p, forward=True, mode=thg.TEMPORAL_SAMPLE_DYNAMIC), p), False)
    assert f[0] == 3 and f[1] is True and f[2] == (-1, 5)
    assert ops._extract_filter((thg.TemporalEdgeFilter((0, 2), p, mode=7), p), False) is None
    with pytest.raises(ValueError, match="homogenous"):
        ops._extract_filter((thg.TemporalEdgeFilter((0, 2), {"a": p}), p), False)
    with pytest.raises(ValueError, match="heterogenous"):
        ops._extract_filter((thg.TemporalEdgeFilter((0, 2), {"a": p}), p), True)
    with pytest.raises(TypeError):
        ops._extract_filter(thg.TemporalEdgeFilter((0, 2), p), False)


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(5, 2, 2)


def test_peer_offsets_reproduce_the_all_to_all_layout():
    """The peer-memory exchange of the partitioned path stores every row where the all-to-all(v) would have delivered
    it; check the offsets against an explicit simulation of both all-to-alls."""
    from tch_geometric.partitioned import peer_offsets
    rng = np.random.default_rng(5)
    for world in (1, 2, 5, 8):
        C = rng.integers(0, 7, (world, world)).tolist()
        # requests: rank q's send buffer is grouped by owner; owner o receives the groups in requester order
        recv = [[(q, o, i) for q in range(world) for i in range(C[q][o])] for o in range(world)]
        # answers: owner o's answers (in received order) go back; requester q receives them in owner order
        back = [[(q, o, i) for o in range(world) for i in range(C[q][o])] for q in range(world)]
        for me in range(world):
            rc, sc, req_row0, ans_row0 = peer_offsets(C, me)
            assert sc == C[me] and rc == [C[q][me] for q in range(world)]
            for o in range(world):      # me as requester: my group for owner o
                for i in range(C[me][o]):
                    assert recv[o][req_row0[o] + i] == (me, o, i)
            for q in range(world):      # me as owner: my answers to requester q
                for i in range(C[q][me]):
                    assert back[q][ans_row0[q] + i] == (q, me, i)


def test_partition_geometry_helpers():
    """Pure host arithmetic of the partitioned plans: column ranges, worst-case frontiers, segment sizes, and the row
    offsets round 1's peer protocol derives from the count matrix."""
    from tch_geometric.partitioned import (cols_per_rank, frontier_caps, partition_bounds, peer_offsets, segment_rows,
                                           SegmentBuffers)
    assert cols_per_rank(10, 4) == 3 and [partition_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert frontier_caps(1024, [15, 10, 5]) == [1024, 15360, 153600]
    # one rank: the segment is the whole frontier (no slack needed); several: slack x the mean load, even, never more than all
    assert segment_rows(256, 153600, 1, 1.5) == 256 * 153600
    s8 = segment_rows(256, 153600, 8, 1.5)
    assert s8 % 2 == 0 and 1.5 * 256 * 153600 / 8 <= s8 <= 1.5 * 256 * 153600 / 8 + 1026
    assert segment_rows(2, 4, 8, 1.5) == 8                       # capped by the frontier itself
    assert segment_rows(256, 153600, 8, 1.5) < 2 ** 26           # fits the slot map's 26-bit row index
    segs, req_words, ans_words = SegmentBuffers.sizes(256, frontier_caps(1024, [15, 10, 5]), [15, 10, 5], 8, 1.5)
    assert req_words == 8 * max(segs) * 2 and ans_words == 8 * max(s * 2 * k for s, k in zip(segs, [15, 10, 5]))
    C = [[1, 2, 3], [4, 5, 6], [7, 8, 9]]                         # C[q][o]: requests of q for owner o
    rc, sc, req_row0, ans_row0 = peer_offsets(C, 1)
    assert sc == [4, 5, 6] and rc == [2, 5, 8]
    assert req_row0 == [1, 2, 3] and ans_row0 == [1, 4, 7]


def test_handles_and_partitioned_entries_validate_on_the_host():
    """The new entry points reject bad arguments before any CUDA call (no GPU needed): graph / plan handles, the batched
    relabel stage and the fixed-segment protocol."""
    lib = N.lib
    h = ctypes.c_void_p(0)
    assert lib.tchgeo_graph_create(0, None, None, None, None, ctypes.addressof(h)) == N.ERR_BAD_ARG
    assert lib.tchgeo_graph_create(1, None, None, None, None, None) == N.ERR_BAD_ARG
    assert lib.tchgeo_graph_derived_bytes(None) == 0
    lib.tchgeo_graph_destroy(None)
    lib.tchgeo_plan_destroy(None)
    assert lib.tchgeo_plan_create(None, ctypes.addressof(h)) == N.ERR_BAD_ARG
    assert lib.tchgeo_plan_enqueue(None, 0, 0, None) == N.ERR_BAD_ARG
    assert lib.tchgeo_plan_collect(None) == N.ERR_BAD_ARG
    assert lib.tchgeo_plan_num_launches(None) == 0
    # relabel geometry: workspace grows with the tree size; absurd sizes are refused
    small = lib.tchgeo_unique_relabel_batched_workspace_bytes(4, 1000, 1)
    big = lib.tchgeo_unique_relabel_batched_workspace_bytes(256, 937984, 1)
    assert 0 < small < big and lib.tchgeo_unique_relabel_batched_workspace_bytes(4, 1 << 40, 1) == 0
    assert lib.tchgeo_unique_relabel_batched(None, 10, None, 4, 2, 20, 1, None, None, None, None, 0, None, None) == N.ERR_BAD_ARG
    # fixed-segment protocol
    assert lib.tchgeo_partf_workspace_bytes(256, 153600) > 0 and lib.tchgeo_partf_workspace_bytes(0, 10) == 0
    assert lib.tchgeo_partf_scatter(None, 0, None, None, 1, 1, 1, 0, 0, 0, 8, None, None, None, None, None, None, None) == N.ERR_BAD_ARG
    assert lib.tchgeo_partf_scatter(None, 0, None, None, 1, 1, 1, 2, 0, 0, 1 << 26, None, None, None, None, None, None, None) == N.ERR_BAD_ARG
    assert b"segment" in lib.tchgeo_last_error()
    assert lib.tchgeo_partf_serve(None, None, None, None, 0, 1, 1, None, None, 8, 0, 65, 0, 0, 0, 1, 0, None, None, None) == N.ERR_BAD_ARG
    assert b"fanout" in lib.tchgeo_last_error()
    assert lib.tchgeo_partf_finish(None, None, 8, 5, None, 1, None, None, 1, 1, None, None, None, None, None, 0, None, None,
                                   None, 0, None, None, 0, None) == N.ERR_BAD_ARG
    assert lib.tchgeo_pack_ragged(None, 4, None, 1, 70000, 4, None, None, None) == N.ERR_BAD_ARG


def test_csx_build_validates_on_the_host():
    """tchgeo_coo_to_csx / tchgeo_ind2ptr refuse bad sizes and NULL pointers before any CUDA call (no GPU needed)."""
    lib = N.lib
    one = ctypes.c_void_p(8)   # never dereferenced: the checks below fail first
    assert lib.tchgeo_coo_to_csx(one, one, -1, 4, 4, 1, one, one, one, one, 1 << 20, None) == N.ERR_BAD_ARG
    assert lib.tchgeo_coo_to_csx(one, one, 10, -4, 4, 1, one, one, one, one, 1 << 20, None) == N.ERR_BAD_ARG
    assert lib.tchgeo_coo_to_csx(one, one, 1 << 32, 4, 4, 1, one, one, one, one, 1 << 20, None) == N.ERR_BAD_ARG
    assert b"2^32" in lib.tchgeo_last_error()
    assert lib.tchgeo_coo_to_csx(one, one, 10, 4, 4, 1, None, one, one, one, 1 << 20, None) == N.ERR_BAD_ARG     # ptrs
    assert lib.tchgeo_coo_to_csx(None, one, 10, 4, 4, 1, one, one, one, one, 1 << 20, None) == N.ERR_BAD_ARG     # row
    assert lib.tchgeo_coo_to_csx(one, one, 10, 4, 4, 1, one, one, one, None, 0, None) == N.ERR_BAD_ARG           # workspace
    assert b"workspace" in lib.tchgeo_last_error()
    assert lib.tchgeo_coo_to_csx(one, one, 10, 1 << 40, 1 << 40, 1, one, one, one, one, 1 << 20, None) == N.ERR_BAD_ARG
    assert b"64 bits" in lib.tchgeo_last_error()
    assert lib.tchgeo_coo_to_csx_workspace_bytes(-1, 4, 4) == 0
    assert lib.tchgeo_ind2ptr(None, 5, 4, one, None) == N.ERR_BAD_ARG
    assert lib.tchgeo_ind2ptr(one, 5, -1, one, None) == N.ERR_BAD_ARG


def test_host_unpack_transport_rebuilds_the_reference_vectors():
    """tchgeo_host_unpack_transport is a pure host function: i32 -> i64 widening and run-length expansion of `cols`,
    any thread count, unaligned destinations, empty batches; inconsistent counts are an error, not garbage."""
    import ctypes
    from tch_geometric import _native as N
    rng = np.random.default_rng(0)
    B = 23
    nn = rng.integers(0, 3000, B)
    nn[3] = 0
    counts = [rng.integers(0, 16, n).astype(np.uint8) for n in nn]
    ne = np.array([int(c.sum()) for c in counts])
    n_off = np.concatenate([[0], np.cumsum(nn)]).astype(np.int64)
    e_off = np.concatenate([[0], np.cumsum(ne)]).astype(np.int64)
    s32 = rng.integers(0, 2**31 - 1, n_off[-1]).astype(np.int32)
    e32 = rng.integers(0, 2**31 - 1, e_off[-1]).astype(np.int32)
    cnt = np.concatenate(counts).astype(np.uint8)
    want_cols = np.concatenate([np.repeat(np.arange(n), c) for n, c in zip(nn, counts)])
    import os
    for shift, simd in ((0, ""), (1, ""), (1, "sse2"), (3, "sse2")):   # aligned / merely 8-byte aligned destinations
        os.environ["TCHGEO_HOST_SIMD"] = simd                          # "sse2": the path of CPUs without AVX-512
        samples = np.zeros(n_off[-1] + 4, np.int64)[shift:shift + n_off[-1]]
        cols = np.zeros(e_off[-1] + 4, np.int64)[shift:shift + e_off[-1]]
        eidx = np.zeros(e_off[-1] + 4, np.int64)[shift:shift + e_off[-1]]
        for threads in (1, 4, 64):
            st = N.lib.tchgeo_host_unpack_transport(s32.ctypes.data, e32.ctypes.data, cnt.ctypes.data, n_off.ctypes.data,
                                                    e_off.ctypes.data, B, samples.ctypes.data, cols.ctypes.data,
                                                    eidx.ctypes.data, threads)
            assert st == 0, N.last_error()
            assert (samples == s32).all() and (eidx == e32).all() and (cols == want_cols).all()
    os.environ.pop("TCHGEO_HOST_SIMD", None)
    cnt[n_off[5]] += 1
    st = N.lib.tchgeo_host_unpack_transport(s32.ctypes.data, e32.ctypes.data, cnt.ctypes.data, n_off.ctypes.data,
                                            e_off.ctypes.data, B, samples.ctypes.data, cols.ctypes.data, eidx.ctypes.data, 2)
    assert st == N.ERR_INTERNAL
