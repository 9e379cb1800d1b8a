"""GPU parity for the partitioned-CSC path: the owner-side serve kernel vs the oracle, and the whole
partitioned sampler (single rank) vs the replicated sampler, bit for bit.  The 2-rank exchange logic is
covered on the CPU (tests/test_partitioned_gloo.py) and on real GPUs by tools/check_partitioned.py."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def thg():
    import tch_geometric
    return tch_geometric


def dev(x, dtype=torch.int64):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).cuda()


@pytest.mark.parametrize("rank,world", [(0, 1), (1, 3), (2, 3)])
@pytest.mark.parametrize("fanout", [1, 5, 15, 70])
def test_serve_kernel_matches_oracle(thg, fakedataset, rank, world, fanout):
    from tch_geometric.partitioned import ColumnPartition, cuda_serve
    ei, n = fakedataset
    ptrs, idx, _ = O.to_csc(ei, n)
    w = np.random.default_rng(3).integers(1, 40, idx.size) / 8.0
    part = ColumnPartition.from_full(dev(ptrs), dev(idx), rank, world, dev(w, torch.float64))
    rng = np.random.default_rng(rank * 10 + fanout)
    m = 700
    ids = rng.integers(part.col_begin, part.col_end, m)
    meta = (rng.integers(0, 50, m) << 32) | rng.integers(0, 100000, m)
    hp = part.ptrs.cpu().numpy()
    hi = part.indices.cpu().numpy()
    for kind, osamp in ((0, None), (1, ("uniform", True)), (2, ("weighted", part.weights.cpu().numpy()))):
        g_ids, g_ptrs = cuda_serve(part, dev(ids), dev(meta), fanout, kind, 777, 0)
        o_ids, o_ptrs = O.serve_requests(hp, hi, part.col_begin, part.edge_base, ids, meta, fanout, sampler=osamp, seed=777)
        assert (g_ids.cpu().numpy() == o_ids).all() and (g_ptrs.cpu().numpy() == o_ptrs).all()
        # answers are global CSC positions of real edges
        ok = o_ptrs >= 0
        assert (idx[o_ptrs[ok]] == o_ids[ok]).all()


def test_partitioned_single_rank_equals_replicated(thg, fakedataset):
    from tch_geometric.partitioned import ColumnPartition, SingleComm
    from partitioned_reference import PartitionedSampler
    ei, n = fakedataset
    ptrs, idx, _ = thg.to_csc(dev(ei), n)
    part = ColumnPartition.from_full(ptrs, idx, 0, 1)
    B, S, fan = 6, 33, [15, 10, 5]
    inputs = dev(np.random.default_rng(1).integers(0, n, (B, S)))
    for sampler in (None, thg.UniformEdgeSampler(True)):
        got = PartitionedSampler(part, fan, sampler, comm=SingleComm()).sample(inputs, seed=9, batch_base=4)
        want = thg.neighbor_sampling_homogenous_batched(ptrs, idx, inputs, fan, sampler, seed=9, batch_base=4)
        for b in range(B):
            for g, x in zip(got[b][:4], want.batch(b)[:4]):
                assert torch.equal(g, x)
            assert list(got[b][4]) == list(want.batch(b)[4])
    with pytest.raises(thg.ReferencePanic):  # out-of-range seed is still an error on the owner
        PartitionedSampler(part, fan, comm=SingleComm()).sample(dev([[n + 5]]), seed=1)


class _FakeWorld:
    """Communicator of a `world`-rank job collapsed onto one GPU: every request comes straight back to the sender,
    grouped by owner exactly as the all-to-all would deliver it to the owners."""

    def __init__(self, world):
        self.rank, self.world = 0, world

    def exchange_rows(self, send_counts, rows, alloc=None):
        sc = send_counts.tolist()
        return sc, sc, rows

    def return_rows(self, rows, n_rows, n_back, send_counts, recv_counts, alloc=None):
        return rows

    def all_gather_int(self, value, device):
        raise AssertionError("the test passes edge_bases explicitly")


@pytest.mark.parametrize("world", [1, 3, 8])
def test_partitioned_plan_equals_replicated(thg, fakedataset, world):
    """The device pipeline (bucket by owner -> serve per owner -> scan + tree layout) against the replicated sampler,
    bit for bit, with `world` column partitions served one after the other on this GPU."""
    from tch_geometric.partitioned import ColumnPartition, PartitionedPlan, SingleComm, serve_rows
    ei, n = fakedataset
    ptrs, idx, _ = thg.to_csc(dev(ei), n)
    w = dev(np.random.default_rng(3).integers(1, 40, idx.numel()) / 8.0, torch.float64)
    parts = [ColumnPartition.from_full(ptrs, idx, r, world, w) for r in range(world)]
    B, S, fan = 7, 33, [15, 10, 5]
    inputs = dev(np.random.default_rng(1).integers(0, n, (B, S)))
    for sampler, kind in ((None, 0), (thg.UniformEdgeSampler(True), 1), (thg.WeightedEdgeSampler(w), 2)):
        plan = None

        def serve(r_req, recv_counts, k, seed, ans, kind=kind):
            o = 0
            for r, c in enumerate(recv_counts):  # owner r answers its group
                if c:
                    serve_rows(parts[r], r_req[o:o + c], c, k, kind, seed, ans[o:o + c], plan.err)
                o += c

        plan = PartitionedPlan(parts[0], B, S, fan, sampler, comm=_FakeWorld(world) if world > 1 else SingleComm(),
                               serve_rows=serve if world > 1 else None, edge_bases=[p.edge_base for p in parts])
        for rep in range(2):  # the plan's buffers are reused across calls
            got = plan.sample(inputs, seed=9 + rep, batch_base=4)
            want = thg.neighbor_sampling_homogenous_batched(ptrs, idx, inputs, fan, sampler, seed=9 + rep, batch_base=4)
            assert (got.samples_len == want.samples_len).all() and (got.edges_len == want.edges_len).all()
            assert (got.layer_offsets == want.layer_offsets).all()
            for b in range(B):
                for g, x in zip(got.batch(b)[:4], want.batch(b)[:4]):
                    assert torch.equal(g, x)
                assert list(got.batch(b)[4]) == list(want.batch(b)[4])


def test_partitioned_plan_errors_and_edge_cases(thg, fakedataset):
    from tch_geometric.partitioned import ColumnPartition, PartitionedPlan, SingleComm
    ei, n = fakedataset
    ptrs, idx, _ = thg.to_csc(dev(ei), n)
    part = ColumnPartition.from_full(ptrs, idx, 0, 1)
    with pytest.raises(thg.ReferencePanic):  # out-of-range seed is still an error on the owner
        PartitionedPlan(part, 1, 1, [3, 2], comm=SingleComm()).sample(dev([[n + 5]]), seed=1)
    # isolated seeds: empty frontiers after the first hop
    ptrs0 = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    got = PartitionedPlan(ColumnPartition.from_full(ptrs0, idx[:0], 0, 1), 2, 3, [4, 4], comm=SingleComm()).sample(
        dev([[0, 1, 2], [3, 4, 5]]), seed=1)
    assert got.samples_len.tolist() == [3, 3] and got.edges_len.tolist() == [0, 0]
    assert got.batch(1)[0].tolist() == [3, 4, 5] and got.batch(1)[4] == [(3, 0, 3), (3, 0, 3)]
    # no hops
    got = PartitionedPlan(part, 2, 3, [], comm=SingleComm()).sample(dev([[0, 1, 2], [3, 4, 5]]), seed=1)
    assert got.samples_len.tolist() == [3, 3] and got.batch(0)[4] == []


@pytest.mark.parametrize("fanout,kind", [(5, 0), (15, 0), (3, 1), (70, 0)])
def test_peer_serve_kernel_places_rows_like_the_answer_all_to_all(thg, fakedataset, fanout, kind):
    """tchgeo_serve_requests_rows_peer with the 'peers' being three local buffers: the answers to requester q's segment
    must land in buffer q from row peer_row0[q] on, and equal what tchgeo_serve_requests_rows writes locally."""
    from tch_geometric import _native as N
    from tch_geometric.partitioned import ColumnPartition, serve_rows
    from tch_geometric.ops import _ptr, _stream
    ei, n = fakedataset
    ptrs, idx, _ = thg.to_csc(dev(ei), n)
    part = ColumnPartition.from_full(ptrs, idx, 1, 3)
    rng = np.random.default_rng(fanout)
    rc = np.array([300, 0, 411], dtype=np.int64)            # requests received from requesters 0, 1, 2
    row0 = np.array([17, 5, 1000], dtype=np.int64)          # where this owner's answers start in each requester's buffer
    m = int(rc.sum())
    req = torch.stack([dev(rng.integers(part.col_begin, part.col_end, m)),
                       dev((rng.integers(0, 50, m) << 32) | rng.integers(0, 100000, m))], dim=1).contiguous()
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    want = torch.full((m, 2 * fanout), -7, dtype=torch.int32, device="cuda")
    serve_rows(part, req, m, fanout, kind, 4242, want, err)
    bufs = [torch.full((2000, 2 * fanout), -7, dtype=torch.int32, device="cuda") for _ in range(3)]
    peer = np.array([b.data_ptr() for b in bufs], dtype=np.uint64)
    N.check(N.lib.tchgeo_serve_requests_rows_peer(_ptr(part.ptrs), _ptr(part.indices), None, part.col_begin,
                                                  part.col_end - part.col_begin, part.indices.numel(), _ptr(req), m, fanout,
                                                  kind, 4242, 0, 3, peer.ctypes.data, rc.ctypes.data, row0.ctypes.data,
                                                  _ptr(err), _stream(req.device)))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    seg = 0
    for q in range(3):
        c, r = int(rc[q]), int(row0[q])
        assert torch.equal(bufs[q][r:r + c], want[seg:seg + c])
        untouched = torch.ones(2000, dtype=torch.bool, device="cuda")
        untouched[r:r + c] = False
        assert (bufs[q][untouched] == -7).all()
        seg += c


def test_peer_request_scatter_places_groups_like_the_request_all_to_all(thg, fakedataset):
    """count_hop + scatter_hop with three local 'owner' buffers: the counts equal begin_hop's, and owner o's buffer
    holds, from row peer_row0[o] on, exactly this rank's group for owner o of the local send buffer."""
    from tch_geometric import _native as N
    from tch_geometric.ops import _ptr, _stream
    ei, n = fakedataset
    B, F, world = 5, 700, 3
    cpr = (n + world - 1) // world
    rng = np.random.default_rng(8)
    samples = dev(rng.integers(0, n, (B, F + 10)))
    fr_end = dev(rng.integers(F - 50, F + 1, B))
    i64 = dict(dtype=torch.int64, device="cuda")
    counts, cursor = torch.zeros(world, **i64), torch.zeros(world, **i64)
    counts2, cursor2 = torch.zeros(world, **i64), torch.zeros(world, **i64)
    req, req2 = torch.zeros((B * F, 2), **i64), torch.zeros((B * F, 2), **i64)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = torch.empty(int(N.lib.tchgeo_part_hop_workspace_bytes(B, F)), dtype=torch.uint8, device="cuda")
    ws2 = torch.empty_like(ws)
    st = _stream(samples.device)
    common = (_ptr(samples), F + 10, None, _ptr(fr_end), B, F, cpr, world)
    N.check(N.lib.tchgeo_part_begin_hop(*common, 3, _ptr(counts2), _ptr(cursor2), _ptr(req2), _ptr(err), _ptr(ws2),
                                        ws2.numel(), st))
    N.check(N.lib.tchgeo_part_count_hop(*common, _ptr(counts), _ptr(cursor), _ptr(err), _ptr(ws), ws.numel(), st))
    torch.cuda.synchronize()
    assert torch.equal(counts, counts2)
    c = counts.tolist()
    assert sum(c) == int(fr_end.sum())
    row0 = np.array([11, 0, 333], dtype=np.int64)
    bufs = [torch.full((B * F + 400, 2), -7, **i64) for _ in range(world)]
    peer = np.array([b.data_ptr() for b in bufs], dtype=np.uint64)
    N.check(N.lib.tchgeo_part_scatter_hop(*common, 3, _ptr(counts), _ptr(cursor), _ptr(req), peer.ctypes.data,
                                          row0.ctypes.data, _ptr(err), _ptr(ws), ws.numel(), st))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    off = 0
    for o in range(world):
        r = int(row0[o])
        assert torch.equal(bufs[o][r:r + c[o]], req[off:off + c[o]])
        assert (bufs[o][:r] == -7).all() and (bufs[o][r + c[o]:] == -7).all()
        # same multiset of requests as the one-call entry point (the order inside a group is arbitrary)
        a = req[off:off + c[o]]
        b = req2[off:off + c[o]]
        ka, kb = a[:, 1].sort().values, b[:, 1].sort().values     # (batch << 32 | pos) is unique per request
        assert torch.equal(ka, kb)
        owner = torch.clamp(a[:, 0] // cpr, max=world - 1)
        assert (owner == o).all()
        off += c[o]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_fixed_segment_protocol_equals_replicated(thg, fakedataset, world):
    """The device-only protocol (csrc/partitioned_fixed.cu: scatter + put -> serve -> one-pass finish, fixed per-pair
    segments, counts on the device) with `world` virtual ranks on this GPU, each sampling its OWN batches from its
    own column range: every rank's result equals the replicated sampler bit for bit, for all three samplers."""
    from tch_geometric.partitioned import (ColumnPartition, PartitionedPlanF, SegmentBuffers, frontier_caps,
                                           sample_virtual_ranks)
    ei, n = fakedataset
    ptrs, idx, _ = thg.to_csc(dev(ei), n)
    w = dev(np.random.default_rng(3).integers(1, 40, idx.numel()) / 8.0, torch.float64)
    parts = [ColumnPartition.from_full(ptrs, idx, r, world, w) for r in range(world)]
    B, S, fan = 5, 33, [15, 10, 5]
    rng = np.random.default_rng(1)
    inputs = [dev(rng.integers(0, n, (B, S))) for _ in range(world)]
    inputs[0][1, :4] = inputs[0][1, 4]                               # duplicated seeds
    bases = [4 + r * B for r in range(world)]
    for sampler in (None, thg.UniformEdgeSampler(True), thg.WeightedEdgeSampler(w)):
        bufs = SegmentBuffers.virtual(B, frontier_caps(S, fan), fan, world, ptrs.device, slack=3.0)
        plans = [PartitionedPlanF(parts[r], B, S, fan, sampler, world=world, rank=r, buffers=bufs[r],
                                  edge_bases=[p.edge_base for p in parts], slack=3.0) for r in range(world)]
        for rep in range(2):                                         # the buffers are reused across steps
            outs = sample_virtual_ranks(plans, inputs, 9 + rep, bases)
            for r in range(world):
                want = thg.neighbor_sampling_homogenous_batched(ptrs, idx, inputs[r], fan, sampler, seed=9 + rep,
                                                                batch_base=bases[r])
                got = outs[r]
                assert (got.samples_len == want.samples_len).all() and (got.edges_len == want.edges_len).all()
                assert (got.layer_offsets == want.layer_offsets).all()
                for b in range(B):
                    for g, x in zip(got.batch(b)[:4], want.batch(b)[:4]):
                        assert torch.equal(g, x)
    thg.clear_caches()


def test_fixed_segment_protocol_errors(thg, fakedataset):
    from tch_geometric.partitioned import (ColumnPartition, PartitionedPlanF, SegmentBuffers, frontier_caps,
                                           sample_virtual_ranks)
    ei, n = fakedataset
    ptrs, idx, _ = thg.to_csc(dev(ei), n)
    world, B, S, fan = 2, 2, 64, [6, 3]
    parts = [ColumnPartition.from_full(ptrs, idx, r, world) for r in range(world)]

    def plans_with(slack):
        bufs = SegmentBuffers.virtual(B, frontier_caps(S, fan), fan, world, ptrs.device, slack=slack)
        return [PartitionedPlanF(parts[r], B, S, fan, None, world=world, rank=r, buffers=bufs[r],
                                 edge_bases=[p.edge_base for p in parts], slack=slack) for r in range(world)]
    # every seed belongs to rank 0's range: rank 0's segment at a tight slack overflows -> capacity error, no corruption
    skew = [dev(np.random.default_rng(r).integers(0, parts[0].col_end, (B, S))) for r in range(world)]
    ps = plans_with(0.6)
    if ps[0].segs[0] < B * S:
        with pytest.raises(MemoryError):
            sample_virtual_ranks(ps, skew, 1, [0, B])
    # an out-of-range seed is reported by the owner it is routed to
    bad = [x.clone() for x in skew]
    bad[1][0, 0] = n + 77
    with pytest.raises((thg.ReferencePanic, MemoryError)):
        sample_virtual_ranks(plans_with(4.0), bad, 1, [0, B])
    thg.clear_caches()


def test_pipelined_groups_equal_replicated(thg, fakedataset):
    """PartitionedPlanGroups: two batch groups on two streams with their own exchange buffers write into one set of
    output buffers; the result equals the replicated sampler (single rank: the exchange is local)."""
    from tch_geometric.partitioned import ColumnPartition, PartitionedPlanGroups
    ei, n = fakedataset
    ptrs, idx, _ = thg.to_csc(dev(ei), n)
    part = ColumnPartition.from_full(ptrs, idx, 0, 1)
    B, S, fan = 7, 33, [15, 10, 5]
    inputs = dev(np.random.default_rng(4).integers(0, n, (B, S)))
    plan = PartitionedPlanGroups(part, B, S, fan, groups=3)
    for rep in range(2):
        got = plan.sample(inputs, seed=21 + rep, batch_base=11)
        want = thg.neighbor_sampling_homogenous_batched(ptrs, idx, inputs, fan, seed=21 + rep, batch_base=11)
        assert (got.samples_len == want.samples_len).all() and (got.layer_offsets == want.layer_offsets).all()
        for b in range(B):
            for g, x in zip(got.batch(b)[:4], want.batch(b)[:4]):
                assert torch.equal(g, x)
    thg.clear_caches()


@pytest.mark.parametrize("groups", [1, 2])
def test_real_ranks_fixed_protocol_equals_replicated(groups):
    """Two real ranks (torchrun, NCCL for the set-up, torch symmetric memory for the exchange buffers): the shipped
    protocol against the replicated sampler on a scaled papers100M-shaped graph, every batch of every rank bit for bit
    (tools/check_partitioned.py; the full-shape runs on 2 and 8 B200s are recorded under profiles/r2_check_partitioned_*).
    Needs two GPUs: skipped on a single-GPU box."""
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + groups), os.path.join(root, "tools", "check_partitioned.py"), "--scale", "0.02",
           "--batches", "8", "--groups", str(groups)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    assert json.loads(line)["ok"] is True
