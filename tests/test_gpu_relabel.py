"""GPU parity: dedup + insertion-order relabel stage vs the oracle (bit-exact)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def thg():
    import tch_geometric
    return tch_geometric


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.int64).cuda()


def test_small_cases(thg):
    nodes, local = thg.unique_relabel(dev([5, 3, 5, 7, 3, 9, 7, 5]), 3)
    assert nodes.tolist() == [5, 3, 5, 7, 9] and local.tolist() == [2, 1, 2, 3, 1, 4, 3, 2]
    nodes, local = thg.unique_relabel(dev([]), 0)
    assert nodes.numel() == 0 and local.numel() == 0
    nodes, local = thg.unique_relabel(dev([4, 4, 4]), 0)
    assert nodes.tolist() == [4] and local.tolist() == [0, 0, 0]


@pytest.mark.parametrize("fan", [[5, 5], [15, 10, 5]])
def test_sampled_trees(thg, fakedataset, fan):
    ei, n = fakedataset
    ptrs, idx, _ = thg.to_csc(dev(ei), n)
    for seeds in (np.arange(64), np.array([7, 7, 2, 7]), np.random.default_rng(0).integers(0, n, 500)):
        samples, rows, cols, eidx, lo = thg.neighbor_sampling_homogenous(ptrs, idx, dev(seeds), fan)
        nodes, local = thg.unique_relabel(samples, len(seeds))
        wn, wl = O.unique_relabel(samples.cpu().numpy(), len(seeds))
        assert (nodes.cpu().numpy() == wn).all() and (local.cpu().numpy() == wl).all()
        # relabeled edges reference the deduplicated node list consistently
        assert torch.equal(nodes[local], samples)
        assert torch.equal(nodes[local[rows]], idx[eidx])
