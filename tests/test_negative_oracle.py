"""CPU: the oracle's negative_sample_neighbors_* (src/algo/negative_sampling.rs) against the reference's own
invariant tests (:147-168, :171-231), the one-answer deterministic case, the HashMap relabel semantic, and
XOSHIRO-vs-COUNTER distributional equivalence."""
import numpy as np

from helpers import chi2_two_sample, has_edge
from oracle import oracle as O


def _hetero_csr(fakehetero):
    counts, edges = fakehetero
    node_types = sorted(counts)
    edge_types = sorted(edges)
    rp, ci, sizes = {}, {}, {}
    for e in edge_types:
        k = O.rel_key(e)
        size = (counts[e[0]], counts[e[2]])
        rp[k], ci[k], _ = O.to_csr(edges[e], size)
        sizes[k] = size
    return counts, node_types, edge_types, rp, ci, sizes


def complete_minus_one(n):
    """row v has every w except v itself (no self loops) and (v + 1) % n: exactly one valid negative per input"""
    src, dst = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    keep = (src != dst) & (dst != (src + 1) % n)
    return O.to_csr(np.stack([src[keep], dst[keep]]), n)[:2]


def test_reference_invariant_homogenous(karate):
    ei, n = karate  # negative_sampling.rs:147-168: all nodes, num_neg 10, try_count 5
    rp, ci, _ = O.to_csr(ei, n)
    for mode in (O.RNG_XOSHIRO, O.RNG_COUNTER):
        samples, rows, cols, count = O.negative_sample_neighbors_homogenous(rp, ci, (n, n), np.arange(n), 10, 5,
                                                                            rng_mode=mode, seed=0)
        assert count == n and (samples[:n] == np.arange(n)).all()
        assert len(np.unique(samples)) == len(samples)          # HashMap: every id appears once
        assert rows.size == cols.size and 0 < rows.size <= n * 10
        assert (np.diff(rows) >= 0).all()                       # generation order
        for i, j in zip(rows, cols):
            v, w = samples[i], samples[j]
            assert not has_edge(rp, ci, v, w) and v != w


def test_reference_invariant_heterogenous(fakehetero):
    counts, node_types, edge_types, rp, ci, sizes = _hetero_csr(fakehetero)
    inputs = {t: np.array([0, 1, 4, 5]) for t in node_types}  # negative_sampling.rs:171-231: num_neg 3, try_count 10
    for mode in (O.RNG_XOSHIRO, O.RNG_COUNTER):
        s, r, c, n = O.negative_sample_neighbors_heterogenous(node_types, edge_types, rp, ci, sizes, inputs, 3, 10,
                                                              False, rng_mode=mode, seed=0)
        assert all(n[t] == 4 for t in node_types)
        total = 0
        for e in edge_types:
            k = O.rel_key(e)
            total += r[k].size
            for i, j in zip(r[k], c[k]):
                v, w = s[e[0]][i], s[e[2]][j]
                assert not has_edge(rp[k], ci[k], v, w) and v != w
        srcs = {e[0] for e in edge_types}
        assert 0 < total <= 3 * 4 * len(srcs)


def test_one_valid_negative_is_deterministic():
    n = 12
    rp, ci = complete_minus_one(n)
    inputs = np.array([3, 7, 3, 0])
    for mode in (O.RNG_XOSHIRO, O.RNG_COUNTER):
        samples, rows, cols, count = O.negative_sample_neighbors_homogenous(rp, ci, (n, n), inputs, 2, 400,
                                                                            rng_mode=mode, seed=5)
        # inputs first (duplicates kept); 4 and 8 are new, 1 is new; a duplicated input maps to its LAST position
        assert samples.tolist() == [3, 7, 3, 0, 4, 8, 1]
        assert rows.tolist() == [0, 0, 1, 1, 2, 2, 3, 3]
        assert cols.tolist() == [4, 4, 5, 5, 4, 4, 6, 6]
        assert count == 4


def test_candidate_equal_to_an_input_maps_to_its_last_slot():
    n = 6
    rp, ci = complete_minus_one(n)
    samples, rows, cols, _ = O.negative_sample_neighbors_homogenous(rp, ci, (n, n), np.array([1, 0, 1]), 1, 300, seed=1)
    assert samples.tolist() == [1, 0, 1, 2]   # negative of 0 is 1 -> maps to the LAST 1 (index 2), nothing appended
    assert rows.tolist() == [0, 1, 2] and cols.tolist() == [3, 2, 3]


def test_exhausted_tries_emit_nothing():
    src, dst = np.meshgrid(np.arange(5), np.arange(5), indexing="ij")
    rp, ci, _ = O.to_csr(np.stack([src.ravel(), dst.ravel()]), 5)  # complete graph with self loops: no negatives
    samples, rows, cols, count = O.negative_sample_neighbors_homogenous(rp, ci, (5, 5), np.arange(5), 3, 7, seed=2)
    assert samples.tolist() == [0, 1, 2, 3, 4] and rows.size == 0 and cols.size == 0 and count == 5
    samples, rows, _, _ = O.negative_sample_neighbors_homogenous(rp, ci, (5, 5), np.arange(5), 0, 7, seed=2)
    assert samples.size == 5 and rows.size == 0


def test_counter_mode_matches_xoshiro_distribution(karate):
    ei, n = karate
    rp, ci, _ = O.to_csr(ei, n)
    inputs = np.tile(np.array([0, 5, 33]), 4000)
    hist = []
    for mode, seed in ((O.RNG_XOSHIRO, 11), (O.RNG_COUNTER, 12)):
        samples, rows, cols, _ = O.negative_sample_neighbors_homogenous(rp, ci, (n, n), inputs, 2, 3, rng_mode=mode, seed=seed)
        hist.append(np.bincount(samples[rows] * n + samples[cols], minlength=n * n))
    assert chi2_two_sample(hist[0], hist[1]) > 0.01
