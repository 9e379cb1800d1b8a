"""CPU: the oracle's tempo_random_walk (src/algo/random_walk.rs:80-158) against the reference's own invariant test
(:332-384), hand-checkable deterministic cases, and XOSHIRO-vs-COUNTER distributional equivalence."""
import numpy as np
import pytest

from helpers import chi2_two_sample
from oracle import oracle as O


def karate_temporal(karate, seed=0):
    ei, n = karate
    rp, ci, _ = O.to_csr(ei, n)
    rng = np.random.default_rng(seed)
    return rp, ci, rng.integers(-1, 5, n), rng.integers(-1, 5, ci.size), n   # gen_range(-1..5), random_walk.rs:340-341


def check_reference_invariants(walks, wts, start, start_ts, window):
    """random_walk.rs:360-383"""
    for i, head in enumerate(start):
        assert walks[i, 0] == head and wts[i, 0] == start_ts[i]
        for t in wts[i]:
            if t == -1 or start_ts[i] == -1:
                continue
            assert start_ts[i] + window[0] <= t < start_ts[i] + window[1]


@pytest.mark.parametrize("mode", [O.RNG_XOSHIRO, O.RNG_COUNTER])
def test_reference_invariants(karate, mode):
    rp, ci, nts, ets, n = karate_temporal(karate)
    start, start_ts = np.array([0, 1, 2, 3]), np.array([0, -1, 2, 3])
    walks, wts = O.tempo_random_walk(rp, ci, nts, ets, start, start_ts, 10, (0, 2), rng_mode=mode, seed=0)
    assert walks.shape == wts.shape == (4, 10)
    check_reference_invariants(walks, wts, start, start_ts, (0, 2))
    assert (walks >= 0).all() and (walks < n).all()   # a walk never stops: it restarts from an earlier position


def test_deterministic_cases():
    # path 0 -> 1 -> 2 (single neighbour each: no draw), 2 has no out-edges: restart draws among earlier positions
    rp, ci, _ = O.to_csr(np.array([[0, 1], [1, 2]]), 3)
    nts, ets = np.array([5, 6, 7]), np.array([-1, 9])
    for mode in (O.RNG_XOSHIRO, O.RNG_COUNTER):
        walks, wts = O.tempo_random_walk(rp, ci, nts, ets, [0], [-1], 3, (0, 1), rng_mode=mode, seed=3)
        assert walks.tolist() == [[0, 1, 2]] and wts.tolist() == [[-1, 6, 9]]   # node ts where the edge ts is -1
        # window [5+1, 5+3) = [6, 8): edge 0->1 has ts 6 (via node 1) and passes, edge 1->2 has ts 9 and fails -> restart
        walks, wts = O.tempo_random_walk(rp, ci, nts, ets, [0], [5], 3, (1, 3), rng_mode=mode, seed=3)
        assert walks[0, :2].tolist() == [0, 1] and wts[0, :2].tolist() == [5, 6]
        assert (walks[0, 2], wts[0, 2]) in ((0, 5), (1, 6))
        # walk_length 1: only the start; isolated start with length 2 restarts on itself
        assert O.tempo_random_walk(rp, ci, nts, ets, [2], [4], 1, (0, 9), rng_mode=mode)[0].tolist() == [[2]]
        walks, wts = O.tempo_random_walk(rp, ci, nts, ets, [2], [4], 2, (0, 9), rng_mode=mode)
        assert walks.tolist() == [[2, 2]] and wts.tolist() == [[4, 4]]
    with pytest.raises(O.OraclePanic):
        O.tempo_random_walk(rp, ci, nts, ets, [0], [0], 0, (0, 1))   # walks_data[i * L] on an empty tensor
    with pytest.raises(O.OraclePanic):
        O.tempo_random_walk(rp, ci, nts, ets, [3], [0], 2, (0, 1))   # start out of range


def test_k1_reservoir_never_keeps_the_first_of_several():
    """quirk Q1 with k = 1: at position 1 the draw U[0,1) is always 0, so the first passing neighbour is always
    replaced when there are at least two."""
    rp, ci, _ = O.to_csr(np.array([[0, 0, 0], [1, 2, 3]]), 4)
    z4, z3 = np.zeros(4, dtype=np.int64), np.zeros(3, dtype=np.int64)
    for mode in (O.RNG_XOSHIRO, O.RNG_COUNTER):
        walks, _ = O.tempo_random_walk(rp, ci, z4, z3, np.zeros(3000, dtype=np.int64), np.zeros(3000, dtype=np.int64), 2,
                                       (0, 1), rng_mode=mode, seed=1)
        c = np.bincount(walks[:, 1], minlength=4)
        assert c[1] == 0 and abs(c[2] - 1500) < 150 and abs(c[3] - 1500) < 150


def test_counter_mode_matches_xoshiro_distribution(karate):
    rp, ci, nts, ets, n = karate_temporal(karate, seed=2)
    start = np.tile(np.arange(n), 300)
    sts = np.tile(np.random.default_rng(3).integers(-1, 4, n), 300)
    h = []
    for mode, seed in ((O.RNG_XOSHIRO, 5), (O.RNG_COUNTER, 6)):
        walks, _ = O.tempo_random_walk(rp, ci, nts, ets, start, sts, 4, (0, 3), rng_mode=mode, seed=seed)
        h.append(np.bincount((walks[:, 0] * n + walks[:, 2]) * n + walks[:, 3], minlength=n ** 3))
    assert chi2_two_sample(h[0], h[1]) > 0.01
