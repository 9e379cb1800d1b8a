"""GPU parity: tempo_random_walk vs the CPU oracle: bit-exact in counter mode, the reference's
invariants, a chi-square test against the sequential (xoshiro) oracle, sharding by walker_base, error behaviour."""
import numpy as np
import pytest
import torch

from helpers import chi2_two_sample
from oracle import oracle as O
from test_tempo_walk_oracle import check_reference_invariants, karate_temporal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def thg():
    import tch_geometric
    return tch_geometric


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.int64).cuda()


def run(thg, rp, ci, nts, ets, start, sts, L, window, **kw):
    w, t = thg.tempo_random_walk(dev(rp), dev(ci), dev(nts), dev(ets), dev(start), dev(sts), L, window, **kw)
    return w.cpu().numpy(), t.cpu().numpy()


@pytest.mark.parametrize("window", [(0, 2), (0, 5), (-3, 1), (2, 2)])
def test_karate_bit_exact_and_invariants(thg, karate, window):
    rp, ci, nts, ets, n = karate_temporal(karate)
    start = np.tile(np.arange(n), 20)
    sts = np.tile(np.random.default_rng(1).integers(-1, 5, n), 20)
    got = run(thg, rp, ci, nts, ets, start, sts, 12, window, seed=7)
    want = O.tempo_random_walk(rp, ci, nts, ets, start, sts, 12, window, seed=7)
    assert (got[0] == want[0]).all() and (got[1] == want[1]).all()
    if window[0] <= 0 < window[1]:  # the reference's invariant presumes the start timestamp lies in its own window
        check_reference_invariants(got[0], got[1], start, sts, window)


def test_fakedataset_heavy_rows_and_sharding(thg, fakedataset):
    ei, n = fakedataset
    rp, ci, _ = O.to_csr(ei, n)
    rng = np.random.default_rng(4)
    nts, ets = rng.integers(-1, 50, n), rng.integers(-1, 50, ci.size)
    # one hub row longer than several 32-lane chunks
    hub = np.stack([np.zeros(200, dtype=np.int64), rng.choice(n, 200, replace=False)])
    rp, ci, perm = O.to_csr(np.concatenate([ei, hub], axis=1), n)
    ets = np.concatenate([ets, rng.integers(-1, 50, 200)])[perm]
    start = np.concatenate([np.zeros(50, dtype=np.int64), np.arange(n)])
    sts = rng.integers(-1, 50, start.size)
    full = run(thg, rp, ci, nts, ets, start, sts, 9, (0, 20), seed=11)
    want = O.tempo_random_walk(rp, ci, nts, ets, start, sts, 9, (0, 20), seed=11)
    assert (full[0] == want[0]).all() and (full[1] == want[1]).all()
    parts = [run(thg, rp, ci, nts, ets, start[a:b], sts[a:b], 9, (0, 20), seed=11, walker_base=a)
             for a, b in ((0, 33), (33, 34), (34, start.size))]
    assert (np.concatenate([p[0] for p in parts]) == full[0]).all()
    assert (np.concatenate([p[1] for p in parts]) == full[1]).all()


def test_warp_per_walker_form_gives_the_same_walks(thg, fakedataset, monkeypatch):
    """The default kernel walks one walker per thread (hubs go to the whole warp); TCHGEO_TEMPO_WALK=warp selects the
    warp-per-walker kernel.  Same draws, same walks -- including walkers on a 300-neighbour hub and on isolated nodes."""
    ei, n = fakedataset
    rng = np.random.default_rng(8)
    hub = np.stack([np.full(300, 3, dtype=np.int64), rng.choice(n, 300, replace=False)])
    rp, ci, perm = O.to_csr(np.concatenate([ei, hub], axis=1), n + 2)     # nodes n, n+1 have no neighbours
    ets = rng.integers(-1, 40, ci.size)
    nts = rng.integers(-1, 40, n + 2)
    start = np.concatenate([np.full(70, 3, dtype=np.int64), np.arange(n + 2)])
    sts = rng.integers(-1, 40, start.size)
    a = run(thg, rp, ci, nts, ets, start, sts, 11, (0, 15), seed=5)
    monkeypatch.setenv("TCHGEO_TEMPO_WALK", "warp")
    b = run(thg, rp, ci, nts, ets, start, sts, 11, (0, 15), seed=5)
    want = O.tempo_random_walk(rp, ci, nts, ets, start, sts, 11, (0, 15), seed=5)
    for x, y, w in zip(a, b, want):
        assert (x == y).all() and (x == w).all()
    # adjacency arrays that are only 8-byte aligned: the thread-per-walker kernel then reads them element by element
    monkeypatch.delenv("TCHGEO_TEMPO_WALK")
    pad = lambda v: torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), dev(v)])[1:]
    w2, t2 = thg.tempo_random_walk(dev(rp), pad(ci), dev(nts), pad(ets), dev(start), dev(sts), 11, (0, 15), seed=5)
    assert (w2.cpu().numpy() == want[0]).all() and (t2.cpu().numpy() == want[1]).all()


def test_distribution_vs_sequential_oracle(thg, karate):
    rp, ci, nts, ets, n = karate_temporal(karate, seed=2)
    start = np.tile(np.arange(n), 300)
    sts = np.tile(np.random.default_rng(3).integers(-1, 4, n), 300)
    g, _ = run(thg, rp, ci, nts, ets, start, sts, 4, (0, 3), seed=21)
    o, _ = O.tempo_random_walk(rp, ci, nts, ets, start, sts, 4, (0, 3), rng_mode=O.RNG_XOSHIRO, seed=22)
    key = lambda w: np.bincount((w[:, 0] * n + w[:, 2]) * n + w[:, 3], minlength=n ** 3)
    assert chi2_two_sample(key(g), key(o)) > 0.01


def test_edge_cases_and_errors(thg):
    rp, ci, _ = O.to_csr(np.array([[0, 1], [1, 2]]), 3)
    nts, ets = np.array([5, 6, 7]), np.array([-1, 9])
    w, t = run(thg, rp, ci, nts, ets, [0], [-1], 3, (0, 1))
    assert w.tolist() == [[0, 1, 2]] and t.tolist() == [[-1, 6, 9]]
    w, t = run(thg, rp, ci, nts, ets, [2], [4], 2, (0, 9))
    assert w.tolist() == [[2, 2]] and t.tolist() == [[4, 4]]
    assert run(thg, rp, ci, nts, ets, [2], [4], 1, (0, 9))[0].tolist() == [[2]]
    assert run(thg, rp, ci, nts, ets, [], [], 5, (0, 9))[0].shape == (0, 5)
    with pytest.raises(thg.ReferencePanic):
        run(thg, rp, ci, nts, ets, [0], [0], 0, (0, 1))
    with pytest.raises(thg.ReferencePanic):
        run(thg, rp, ci, nts, ets, [3], [0], 2, (0, 1))
