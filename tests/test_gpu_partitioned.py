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
