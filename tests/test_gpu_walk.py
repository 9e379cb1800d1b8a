"""GPU parity: node2vec random_walk vs the CPU oracle (bit-exact in counter mode) + invariants."""
import numpy as np
import pytest
import torch

from helpers import chi2_two_sample, has_edge
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def thg():
    import tch_geometric
    return tch_geometric


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.int64).cuda()


@pytest.mark.parametrize("p,q", [(1.0, 1.5), (1.0, 0.5), (0.25, 4.0), (2.0, 1.0), (1.0, 1.0)])
def test_karate_bit_exact_and_valid(thg, karate, p, q):
    ei, n = karate
    rp, ci, _ = thg.to_csr(dev(ei), n)
    hrp, hci = rp.cpu().numpy(), ci.cpu().numpy()
    start = np.array([0, 1, 2, 3])
    walks, att = thg.random_walk(rp, ci, dev(start), 10, p, q, seed=3, return_attempts=True)
    walks = walks.cpu().numpy()
    want, watt = O.random_walk(hrp, hci, start, 10, p, q, seed=3, return_attempts=True)
    assert (walks == want).all() and att == watt
    # the reference's own invariants, random_walk.rs:322-330
    assert (walks[:, 0] == start).all()
    for w in walks:
        for a, b in zip(w[:-1], w[1:]):
            assert has_edge(hrp, hci, a, b)


def test_fakedataset_all_nodes(thg, fakedataset):
    ei, n = fakedataset
    rp, ci, _ = thg.to_csr(dev(ei), n)
    hrp, hci = rp.cpu().numpy(), ci.cpu().numpy()
    start = np.tile(np.arange(n), 3)
    for L in (1, 15, 16, 17, 80):  # around the 16-column staging chunk
        walks = thg.random_walk(rp, ci, dev(start), L, 1.0, 0.5, seed=11).cpu().numpy()
        assert walks.shape == (start.size, L + 1)
        assert (walks == O.random_walk(hrp, hci, start, L, 1.0, 0.5, seed=11)).all()


def test_dead_ends_and_sharding(thg):
    # directed path with a sink plus an isolated node: rows stay -1 after the walk stops (random_walk.rs:45-47)
    ei = np.array([[0, 1, 1, 3], [1, 2, 3, 1]])
    rp, ci, _ = thg.to_csr(dev(ei), 5)
    hrp, hci = rp.cpu().numpy(), ci.cpu().numpy()
    start = np.array([0, 2, 4, 1, 3] * 40)
    full = thg.random_walk(rp, ci, dev(start), 6, 1.0, 2.0, seed=5).cpu().numpy()
    assert (full == O.random_walk(hrp, hci, start, 6, 1.0, 2.0, seed=5)).all()
    assert (full[1] == [2, -1, -1, -1, -1, -1, -1]).all() and (full[2] == [4, -1, -1, -1, -1, -1, -1]).all()
    # walker_base makes sharded launches reproduce the single launch (multi-GPU walker sharding)
    parts = [thg.random_walk(rp, ci, dev(start[a:b]), 6, 1.0, 2.0, seed=5, walker_base=a).cpu().numpy()
             for a, b in ((0, 70), (70, 71), (71, 200))]
    assert (np.concatenate(parts) == full).all()
    assert thg.random_walk(rp, ci, dev(np.zeros(0, dtype=np.int64)), 6, 1.0, 2.0).shape == (0, 7)
    assert thg.random_walk(rp, ci, dev([1]), 0, 1.0, 2.0).tolist() == [[1]]


def test_transition_law_vs_sequential_oracle(thg, karate):
    ei, n = karate
    rp, ci, _ = thg.to_csr(dev(ei), n)
    hrp, hci = rp.cpu().numpy(), ci.cpu().numpy()
    start = np.tile(np.arange(n), 600)
    g = thg.random_walk(rp, ci, dev(start), 3, 0.5, 2.0, seed=17).cpu().numpy()
    o = O.random_walk(hrp, hci, start, 3, 0.5, 2.0, rng_mode=O.RNG_XOSHIRO, seed=23)
    key = lambda w: np.bincount((w[:, 1] * n + w[:, 2]) * n + w[:, 3], minlength=n ** 3)
    assert chi2_two_sample(key(g), key(o)) > 0.01


def test_errors(thg, karate):
    ei, n = karate
    rp, ci, _ = thg.to_csr(dev(ei), n)
    with pytest.raises(ValueError):
        thg.random_walk(rp, ci, dev([0]), 5, 0.0, 1.0)   # "p or q may not be 0 or nan"
    with pytest.raises(thg.ReferencePanic):
        thg.random_walk(rp, ci, dev([34]), 5, 1.0, 1.0)  # start out of range
    with pytest.raises(ValueError):
        thg.random_walk(rp.cpu(), ci, dev([0]), 5, 1.0, 1.0)
