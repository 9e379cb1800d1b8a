"""GPU parity for the temporal filter: bit-exact vs the oracle for every filter mode,
direction and sampler, homogeneous and heterogeneous, plus the reference's own window invariants."""
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


def next_seed(thg, state):
    thg.rng_reseed(state)
    return thg.ops.splitmix64(state)[1]


MODES = [(0, False, (0, 2)), (1, False, (0, 2)), (1, True, (0, 3)), (2, False, (-1, 2)), (2, True, (0, 1))]


@pytest.mark.parametrize("mode,forward,window", MODES)
@pytest.mark.parametrize("which", ["karate", "fake"])
def test_homogenous_filter_bit_exact(thg, karate, fakedataset, mode, forward, window, which):
    ei, n = karate if which == "karate" else fakedataset
    ptrs, idx, _ = thg.to_csc(dev(ei), n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    rng = np.random.default_rng(mode * 7 + forward)
    ts = rng.integers(0, 5, hi.size)
    inputs = rng.integers(0, n, 40 if which == "karate" else 300)
    st = rng.integers(0, 5, inputs.size)
    w = rng.integers(1, 40, hi.size) / 8.0
    fan = [4, 3] if which == "karate" else [15, 10, 5]
    for sampler, osamp in ((None, None), (thg.UniformEdgeSampler(True), ("uniform", True)),
                           (thg.WeightedEdgeSampler(dev(w, torch.float64)), ("weighted", w))):
        seed = next_seed(thg, 11 + mode)
        flt = (thg.TemporalEdgeFilter(window, dev(ts), forward, mode), dev(st))
        got = thg.neighbor_sampling_homogenous(ptrs, idx, dev(inputs), fan, sampler, flt)
        want = O.neighbor_sampling_homogenous(hp, hi, inputs, fan, sampler=osamp, seed=seed,
                                              filter=dict(mode=mode, forward=forward, window=window, timestamps=ts, inputs_state=st))
        for g, x in zip(got[:4], want[:4]):
            assert g.numel() == x.size and (g.cpu().numpy() == x).all()
        assert list(got[4]) == list(want[4])
        e = got[3].cpu().numpy()
        if mode == 0 and len(e):  # the reference's static-window invariant, neighbor_sampling.rs:528-536
            assert ((ts[e] >= window[0]) & (ts[e] <= window[1])).all()


def test_unknown_mode_is_identity_and_errors(thg, karate):
    ei, n = karate
    ptrs, idx, _ = thg.to_csc(dev(ei), n)
    ts = dev(np.zeros(idx.numel(), dtype=np.int64))
    inp = dev([0, 1, 2])
    thg.rng_reseed(3)
    a = thg.neighbor_sampling_homogenous(ptrs, idx, inp, [3, 2])
    thg.rng_reseed(3)
    b = thg.neighbor_sampling_homogenous(ptrs, idx, inp, [3, 2], None, (thg.TemporalEdgeFilter((5, 6), ts, mode=9), inp))
    assert torch.equal(a[3], b[3])                      # python.rs:249: falls through to IdentityFilter
    with pytest.raises(thg.ReferencePanic):             # state vector shorter than inputs (quirk Q10)
        thg.neighbor_sampling_homogenous(ptrs, idx, inp, [3], None, (thg.TemporalEdgeFilter((0, 1), ts), dev([0])))
    out = thg.neighbor_sampling_homogenous(ptrs, idx, inp, [3, 3], None, (thg.TemporalEdgeFilter((7, 9), ts), inp))
    assert out[1].numel() == 0 and out[4] == [(3, 0, 3), (3, 0, 3)]  # everything filtered out


def test_heterogenous_filter_bit_exact(thg, fakehetero):
    counts, edges = fakehetero
    node_types, edge_types = sorted(counts), sorted(edges)
    cp, ri, hcp, hri, ts, hts = {}, {}, {}, {}, {}, {}
    rng = np.random.default_rng(5)
    for et in edge_types:
        k = thg.rel_key(et)
        p, i, _ = thg.to_csc(dev(edges[et]), (counts[et[0]], counts[et[2]]))
        cp[k], ri[k], hcp[k], hri[k] = p, i, p.cpu().numpy(), i.cpu().numpy()
        hts[k] = rng.integers(0, 6, hri[k].size)
        ts[k] = dev(hts[k])
    inputs = {t: rng.integers(0, 800, 9) for t in node_types}
    states = {t: rng.integers(0, 6, 9) for t in node_types}
    nn = {thg.rel_key(et): [4, 3] for et in edge_types}
    for mode, forward, window in ((0, False, (1, 3)), (1, False, (0, 2)), (2, True, (0, 2))):
        seed = next_seed(thg, 40 + mode)
        flt = (thg.TemporalEdgeFilter(window, ts, forward, mode), {t: dev(v) for t, v in states.items()})
        got = thg.neighbor_sampling_heterogenous(node_types, edge_types, cp, ri, {t: dev(v) for t, v in inputs.items()},
                                                 nn, 2, None, flt)
        want = O.neighbor_sampling_heterogenous(node_types, edge_types, hcp, hri, inputs, nn, 2, seed=seed,
                                                filter=dict(mode=mode, forward=forward, window=window, timestamps=hts,
                                                            inputs_state=states))
        for t in node_types:
            assert (got[0][t].cpu().numpy() == want[0][t]).all()
        for k in want[1]:
            assert (got[1][k].cpu().numpy() == want[1][k]).all() and (got[2][k].cpu().numpy() == want[2][k]).all()
            assert (got[3][k].cpu().numpy() == want[3][k]).all() and list(got[4][k]) == list(want[4][k])
