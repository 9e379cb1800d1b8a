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
    from tch_geometric.partitioned import ColumnPartition, PartitionedSampler, SingleComm
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
