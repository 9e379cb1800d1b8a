"""GPU parity: negative_sample_neighbors_homogenous / _heterogenous against the CPU oracle:
bit-exact in counter mode, the reference's invariants, the one-answer case against the sequential (xoshiro) oracle,
and a chi-square test of the chosen negatives against the sequential oracle."""
import numpy as np
import pytest
import torch

from helpers import chi2_two_sample, has_edge
from oracle import oracle as O
from test_negative_oracle import _hetero_csr, complete_minus_one

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def thg():
    import tch_geometric
    return tch_geometric


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.int64).cuda()


def host(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("num_neg,try_count", [(10, 5), (1, 1), (3, 40)])
def test_homogenous_bit_exact_and_invariants(thg, karate, num_neg, try_count):
    ei, n = karate
    rp, ci, _ = O.to_csr(ei, n)
    inputs = np.arange(n)
    got = thg.negative_sample_neighbors_homogenous(dev(rp), dev(ci), (n, n), dev(inputs), num_neg, try_count, seed=9)
    want = O.negative_sample_neighbors_homogenous(rp, ci, (n, n), inputs, num_neg, try_count, seed=9)
    for g, w in zip(got[:3], want[:3]):
        assert (host(g) == w).all()
    assert got[3] == want[3] == n
    samples, rows, cols = (host(x) for x in got[:3])
    for i, j in zip(rows, cols):  # negative_sampling.rs:163-166
        assert not has_edge(rp, ci, samples[i], samples[j]) and samples[i] != samples[j]


def test_homogenous_larger_graph_with_duplicate_inputs(thg, fakedataset):
    ei, n = fakedataset
    rp, ci, _ = O.to_csr(ei, n)
    inputs = np.random.default_rng(3).integers(0, n, 5000)  # many duplicated inputs
    got = thg.negative_sample_neighbors_homogenous(dev(rp), dev(ci), (n, n), dev(inputs), 7, 4, seed=21)
    want = O.negative_sample_neighbors_homogenous(rp, ci, (n, n), inputs, 7, 4, seed=21)
    for g, w in zip(got[:3], want[:3]):
        assert (host(g) == w).all()


def test_one_valid_negative_matches_the_sequential_oracle(thg):
    n = 12
    rp, ci = complete_minus_one(n)
    inputs = np.array([3, 7, 3, 0])
    got = thg.negative_sample_neighbors_homogenous(dev(rp), dev(ci), (n, n), dev(inputs), 2, 400, seed=1)
    want = O.negative_sample_neighbors_homogenous(rp, ci, (n, n), inputs, 2, 400, rng_mode=O.RNG_XOSHIRO, seed=77)
    for g, w in zip(got[:3], want[:3]):
        assert (host(g) == w).all()
    assert host(got[0]).tolist() == [3, 7, 3, 0, 4, 8, 1]


def test_exhausted_tries_and_empty_inputs(thg):
    src, dst = np.meshgrid(np.arange(5), np.arange(5), indexing="ij")
    rp, ci, _ = O.to_csr(np.stack([src.ravel(), dst.ravel()]), 5)
    samples, rows, cols, count = thg.negative_sample_neighbors_homogenous(dev(rp), dev(ci), (5, 5), dev(np.arange(5)), 3, 7)
    assert host(samples).tolist() == [0, 1, 2, 3, 4] and rows.numel() == 0 and cols.numel() == 0 and count == 5
    samples, rows, cols, count = thg.negative_sample_neighbors_homogenous(dev(rp), dev(ci), (5, 5), dev([]), 3, 7)
    assert samples.numel() == 0 and rows.numel() == 0 and count == 0
    samples, rows, cols, count = thg.negative_sample_neighbors_homogenous(dev(rp), dev(ci), (5, 5), dev([1, 2]), 0, 7)
    assert host(samples).tolist() == [1, 2] and rows.numel() == 0
    with pytest.raises(thg.ReferencePanic):  # ptrs[v + 1] out of bounds in the reference
        thg.negative_sample_neighbors_homogenous(dev(rp), dev(ci), (5, 5), dev([5]), 1, 1)


@pytest.mark.parametrize("inbound", [False, True])
def test_heterogenous_bit_exact_and_invariants(thg, fakehetero, inbound):
    counts, node_types, edge_types, rp, ci, sizes = _hetero_csr(fakehetero)
    if inbound:  # has_edge(w, v) indexes the rows with the candidate: keep the relations whose CSR is square
        edge_types = [e for e in edge_types if e[0] == e[2]]
        node_types = sorted({e[0] for e in edge_types})
        assert edge_types
    rels = [O.rel_key(e) for e in edge_types]
    rp, ci, sizes = ({k: d[k] for k in rels} for d in (rp, ci, sizes))
    inputs = {t: np.array([0, 1, 4, 5, 4]) for t in node_types}
    got = thg.negative_sample_neighbors_heterogenous(
        node_types, edge_types, {k: dev(v) for k, v in rp.items()}, {k: dev(v) for k, v in ci.items()}, sizes,
        {t: dev(v) for t, v in inputs.items()}, 3, 10, inbound, seed=31)
    want = O.negative_sample_neighbors_heterogenous(node_types, edge_types, rp, ci, sizes, inputs, 3, 10, inbound, seed=31)
    assert got[3] == want[3]
    for t in node_types:
        assert (host(got[0][t]) == want[0][t]).all()
    for e in edge_types:
        k = O.rel_key(e)
        assert (host(got[1][k]) == want[1][k]).all() and (host(got[2][k]) == want[2][k]).all()
        for i, j in zip(want[1][k], want[2][k]):  # negative_sampling.rs:221-229
            v, w = want[0][e[0]][i], want[0][e[2]][j]
            assert not (has_edge(rp[k], ci[k], w, v) if inbound else has_edge(rp[k], ci[k], v, w)) and v != w


def test_distribution_matches_the_sequential_oracle(thg, karate):
    ei, n = karate
    rp, ci, _ = O.to_csr(ei, n)
    inputs = np.tile(np.array([0, 5, 33]), 4000)
    s, r, c, _ = thg.negative_sample_neighbors_homogenous(dev(rp), dev(ci), (n, n), dev(inputs), 2, 3, seed=41)
    s, r, c = host(s), host(r), host(c)
    os_, or_, oc, _ = O.negative_sample_neighbors_homogenous(rp, ci, (n, n), inputs, 2, 3, rng_mode=O.RNG_XOSHIRO, seed=42)
    assert chi2_two_sample(np.bincount(s[r] * n + s[c], minlength=n * n),
                           np.bincount(os_[or_] * n + os_[oc], minlength=n * n)) > 0.01
