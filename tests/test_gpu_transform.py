"""GPU parity: per-column CSC transforms (src/data/transform.rs) vs the reference's KATs and the
CPU oracle, and the weighted sampler reading the prefix sums (bit-exact for arbitrary weights)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

KAT_PTRS = [0, 0, 0, 0, 3, 5, 5, 5, 7, 9]
KAT_W = [9.0, 5.0, 8.0, 9.0, 10.0, 11.0, 1.0, 1.5]


@pytest.fixture(scope="module")
def thg():
    import tch_geometric
    return tch_geometric


def i64(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.int64).cuda()


def f64(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64).cuda()


def test_cumsum_kat(thg):
    data = f64(KAT_W)  # transform.rs:85-97 (the last column pointer lies past the data: clamped like Tensor::slice)
    thg.csc_edge_cumsum(i64(KAT_PTRS), data)
    assert data.tolist() == [9.0, 14.0, 22.0, 9.0, 19.0, 11.0, 12.0, 1.5]


def test_sort_edges_kat(thg):
    got = thg.csc_sort_edges(i64(KAT_PTRS), i64(np.arange(8)), f64(KAT_W), False)  # transform.rs:69-82
    assert got.tolist() == [1, 2, 0, 3, 4, 6, 5, 7]
    got = thg.csc_sort_edges(i64([0, 3, 5]), i64([10, 11, 12, 13, 14]), f64([1.0, 3.0, 2.0, 5.0, 5.0]), True)
    assert got.tolist() == [11, 12, 10, 13, 14]


def _random_csc(rng, n_cols, max_deg):
    deg = rng.integers(0, max_deg, n_cols)
    deg[rng.integers(0, n_cols, 3)] = max_deg * 40  # a few heavy columns
    ptrs = np.concatenate([[0], np.cumsum(deg)])
    return ptrs, int(ptrs[-1])


def test_cumsum_matches_serial_oracle_bit_for_bit(thg):
    rng = np.random.default_rng(5)
    ptrs, nnz = _random_csc(rng, 5000, 60)
    w = rng.uniform(0.2, 5.0, nnz) * 10.0 ** rng.integers(-6, 6, nnz)  # sums that depend on the order of addition
    data = f64(w)
    thg.csc_edge_cumsum(i64(ptrs), data)
    want = O.csc_edge_cumsum(ptrs, w)
    assert (data.cpu().numpy() == want).all()
    # empty inputs
    thg.csc_edge_cumsum(i64([0]), f64([]))
    thg.csc_edge_cumsum(i64([0, 0, 0]), f64([]))


@pytest.mark.parametrize("descending", [False, True])
def test_sort_edges_matches_oracle(thg, descending):
    rng = np.random.default_rng(6)
    ptrs, nnz = _random_csc(rng, 3000, 50)
    w = np.round(rng.uniform(0, 4, nnz), 1)  # many ties
    perm = rng.permutation(nnz)
    got = thg.csc_sort_edges(i64(ptrs), i64(perm), f64(w), descending).cpu().numpy()
    assert (got == O.csc_sort_edges(ptrs, perm, w, descending)).all()
    assert thg.csc_sort_edges(i64([0, 0]), i64([]), f64([]), descending).numel() == 0


@pytest.mark.parametrize("cumsum", ["1", "0"])
def test_weighted_sampler_arbitrary_weights(thg, fakedataset, monkeypatch, cumsum):
    """With the prefix sums the weighted sampler compares exactly the reference's f64 values, so the output equals the
    serial oracle bit for bit for weights whose sums depend on the order of addition.  The scan path (cumsum off)
    is checked on the same input for structure only."""
    monkeypatch.setenv("TCHGEO_WEIGHT_CUMSUM", cumsum)
    thg.clear_caches()
    ei, n = fakedataset
    ptrs, idx, _ = thg.to_csc(i64(ei), n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    rng = np.random.default_rng(8)
    w = rng.uniform(0.2, 5.0, hi.size) * 10.0 ** rng.integers(-8, 8, hi.size)
    seeds = rng.integers(0, n, 200)
    thg.rng_reseed(77)
    seed = thg.ops.splitmix64(77)[1]
    got = thg.neighbor_sampling_homogenous(ptrs, idx, i64(seeds), [6, 4], thg.WeightedEdgeSampler(f64(w)))
    want = O.neighbor_sampling_homogenous(hp, hi, seeds, [6, 4], sampler=("weighted", w), seed=seed)
    if cumsum == "1":
        for g, x in zip(got[:4], want[:4]):
            assert (g.cpu().numpy() == x).all()
        assert list(got[4]) == list(want[4])
    else:
        s, r, c, e = (t.cpu().numpy() for t in got[:4])
        assert s.size == want[0].size and (hi[e] == s[r]).all()
    thg.clear_caches()


def test_weighted_sampler_heavy_columns_bit_exact(thg, fakedataset):
    """columns far above fanout + 128 take the warp-per-node path of the flattened weighted draw loop"""
    ei, n = fakedataset
    rng = np.random.default_rng(12)
    hubs = [np.stack([rng.choice(n, d, replace=False), np.full(d, c)]) for c, d in ((3, 700), (9, 140), (11, 133))]
    ptrs, idx, _ = thg.to_csc(i64(np.concatenate([ei] + hubs, axis=1)), n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    w = rng.uniform(0.2, 5.0, hi.size) * 10.0 ** rng.integers(-3, 3, hi.size)
    seeds = np.concatenate([[3, 9, 11, 3], rng.integers(0, n, 300)])
    for fan in ([5], [12, 3]):
        thg.rng_reseed(5)
        seed = thg.ops.splitmix64(5)[1]
        got = thg.neighbor_sampling_homogenous(ptrs, idx, i64(seeds), fan, thg.WeightedEdgeSampler(f64(w)))
        want = O.neighbor_sampling_homogenous(hp, hi, seeds, fan, sampler=("weighted", w), seed=seed)
        for g, x in zip(got[:4], want[:4]):
            assert (g.cpu().numpy() == x).all()
