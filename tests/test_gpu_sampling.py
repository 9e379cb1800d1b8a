"""GPU parity tests for neighbor sampling: CUDA path (through the C ABI) vs the CPU oracle.

The oracle's counter mode shares the Philox counter layout with the kernels, so every comparison
here is BIT-EXACT for all five outputs, in the deterministic regime and in the stochastic one.
Statistical tests against closed-form marginals and the sequential-RNG oracle are on top.
"""
import numpy as np
import pytest
import torch

from helpers import (chi2_pvalue, chi2_two_sample, full_neighborhood_tree, reservoir_inclusion,
                     validate_neighbor_samples, validate_tree_identities)
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


def graph(thg, ei, n):
    ptrs, idx, perm = thg.to_csc(dev(ei), n)
    return ptrs, idx, perm


def run_homo(thg, ptrs, idx, inputs, fan, sampler=None, state=1):
    seed = next_seed(thg, state)
    out = thg.neighbor_sampling_homogenous(ptrs, idx, dev(inputs), fan, sampler)
    return [t.cpu().numpy() for t in out[:4]] + [out[4]], seed


def oracle_sampler(sampler):
    if sampler is None:
        return None
    if hasattr(sampler, "with_replacement"):
        return ("uniform", sampler.with_replacement)
    return ("weighted", sampler.weights.cpu().numpy())


def assert_same(got, want):
    for name, g, w in zip(("samples", "rows", "cols", "edge_index"), got[:4], want[:4]):
        assert g.shape == w.shape, name
        assert (g == w).all(), name
    assert list(got[4]) == list(want[4])


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fan", [[5, 5], [4, 3], [17, 17], [1], [3, 2, 2]])
@pytest.mark.parametrize("inputs", [[0, 1, 4, 5], list(range(34)), [33, 33, 0]])
def test_karate_bit_exact(thg, karate, fan, inputs):
    ei, n = karate
    ptrs, idx, _ = graph(thg, ei, n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    for sampler in (None, thg.UniformEdgeSampler(True), thg.UniformEdgeSampler(False)):
        got, seed = run_homo(thg, ptrs, idx, inputs, fan, sampler, state=len(inputs) * 100 + sum(fan))
        want = O.neighbor_sampling_homogenous(hp, hi, inputs, fan, sampler=oracle_sampler(sampler), seed=seed)
        assert_same(got, want)
        validate_neighbor_samples(hp, hi, got[1], got[2], got[0], got[0], got[4], fan)
        validate_tree_identities(hp, hi, np.asarray(inputs), *got, fan,
                                 replace=bool(getattr(sampler, "with_replacement", False)))


def test_deterministic_regime(thg, karate):
    ei, n = karate
    ptrs, idx, _ = graph(thg, ei, n)
    want = full_neighborhood_tree(ptrs.cpu().numpy(), idx.cpu().numpy(), np.arange(34), 2)
    got, _ = run_homo(thg, ptrs, idx, np.arange(34), [17, 17])
    assert_same(got, want)
    got, _ = run_homo(thg, ptrs, idx, np.arange(34), [40, 100])
    assert_same(got, want)


@pytest.mark.parametrize("fan", [[15, 10, 5], [2, 2], [40, 3], [300]])
def test_fakedataset_bit_exact(thg, fakedataset, fan):
    ei, n = fakedataset
    ptrs, idx, _ = graph(thg, ei, n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    inputs = np.random.default_rng(0).integers(0, n, 200)
    w = torch.as_tensor(np.random.default_rng(1).uniform(0.2, 5.0, hi.size)).cuda()
    for sampler in (None, thg.UniformEdgeSampler(True), thg.WeightedEdgeSampler(w)):
        got, seed = run_homo(thg, ptrs, idx, inputs, fan, sampler, state=7)
        want = O.neighbor_sampling_homogenous(hp, hi, inputs, fan, sampler=oracle_sampler(sampler), seed=seed)
        assert_same(got, want)


def test_weighted_dyadic_weights_bit_exact(thg, karate):
    """weights that are multiples of 1/8 sum exactly in any order, so even the f64 path is bit-exact."""
    ei, n = karate
    ptrs, idx, _ = graph(thg, ei, n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    w = np.random.default_rng(2).integers(1, 40, hi.size) / 8.0
    sampler = thg.WeightedEdgeSampler(torch.as_tensor(w).cuda())
    for state in range(5):
        got, seed = run_homo(thg, ptrs, idx, np.arange(34), [4, 3], sampler, state=state)
        want = O.neighbor_sampling_homogenous(hp, hi, np.arange(34), [4, 3], sampler=("weighted", w), seed=seed)
        assert_same(got, want)


def _power_law_graph(n, seed, max_deg):
    rng = np.random.default_rng(seed)
    deg = np.minimum((rng.pareto(1.2, n) * 4).astype(np.int64), max_deg)
    deg[rng.integers(0, n, n // 10)] = 0
    deg[0] = max_deg  # one hub much larger than a tile
    col = np.repeat(np.arange(n), deg)
    row = rng.integers(0, n, col.size)
    ei = np.unique(np.stack([row, col]), axis=1)
    return ei


def test_hub_and_zero_degree_nodes(thg):
    """hub column (deg >> 256*4 draws) exercises the flattened draw blocks; zero-degree nodes yield nothing."""
    n = 5000
    ei = _power_law_graph(n, 3, 4000)
    ptrs, idx, _ = graph(thg, ei, n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    assert (hp == O.to_csc(ei, n)[0]).all()
    inputs = np.concatenate([[0, 0, 0], np.random.default_rng(4).integers(0, n, 700)])
    for sampler, fan in ((None, [15, 10, 5]), (thg.UniformEdgeSampler(True), [6, 3]), (None, [600, 2])):
        got, seed = run_homo(thg, ptrs, idx, inputs, fan, sampler, state=9)
        want = O.neighbor_sampling_homogenous(hp, hi, inputs, fan, sampler=oracle_sampler(sampler), seed=seed)
        assert_same(got, want)


def test_batched_equals_per_batch_calls(thg, fakedataset):
    ei, n = fakedataset
    ptrs, idx, _ = graph(thg, ei, n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    B, S, fan = 12, 37, [7, 4, 3]
    inputs = np.random.default_rng(5).integers(0, n, (B, S))
    res = thg.neighbor_sampling_homogenous_batched(ptrs, idx, dev(inputs), fan, seed=1234, batch_base=5)
    assert len(res) == B
    for b in range(B):
        out = res.batch(b)
        got = [t.cpu().numpy() for t in out[:4]] + [out[4]]
        want = O.neighbor_sampling_homogenous(hp, hi, inputs[b], fan, seed=1234, batch=5 + b)
        assert_same(got, want)
    # a reusable plan gives the same answer and can be re-run
    plan = thg.HomogenousSampler(ptrs, idx, B, S, fan)
    for _ in range(2):
        res2 = plan.sample(dev(inputs), seed=1234, batch_base=5)
        assert (res2.samples_len == res.samples_len).all()
        for b in (0, B - 1):
            for x, y in zip(res2.batch(b)[:4], res.batch(b)[:4]):
                assert torch.equal(x, y)


def test_empty_and_degenerate_inputs(thg, karate):
    ei, n = karate
    ptrs, idx, _ = graph(thg, ei, n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    got, seed = run_homo(thg, ptrs, idx, np.zeros(0, dtype=np.int64), [3, 2])
    assert [len(x) for x in got[:4]] == [0, 0, 0, 0] and got[4] == [(0, 0, 0), (0, 0, 0)]
    got, seed = run_homo(thg, ptrs, idx, [5], [])
    assert got[0].tolist() == [5] and got[4] == []
    # graph without edges
    p0 = torch.zeros(11, dtype=torch.int64).cuda()
    i0 = torch.zeros(0, dtype=torch.int64).cuda()
    got, seed = run_homo(thg, p0, i0, [1, 2], [3, 3])
    assert got[0].tolist() == [1, 2] and len(got[1]) == 0 and got[4] == [(2, 0, 2), (2, 0, 2)]
    # with replacement, fanout 0 is legal and yields nothing
    got, seed = run_homo(thg, ptrs, idx, [0, 1], [0], thg.UniformEdgeSampler(True))
    assert len(got[1]) == 0


def test_reference_panics_are_errors(thg, karate):
    ei, n = karate
    ptrs, idx, _ = graph(thg, ei, n)
    with pytest.raises(thg.ReferencePanic):  # seed out of range (quirk Q10)
        thg.neighbor_sampling_homogenous(ptrs, idx, dev([34]), [2])
    with pytest.raises(thg.ReferencePanic):
        thg.neighbor_sampling_homogenous(ptrs, idx, dev([-1]), [2])
    with pytest.raises(thg.ReferencePanic):  # gen_range(0..0), sampling.rs:19
        thg.neighbor_sampling_homogenous(ptrs, idx, dev([0]), [0])
    with pytest.raises(ValueError):          # wrong device: TensorConversionError -> PyValueError
        thg.neighbor_sampling_homogenous(ptrs.cpu(), idx, dev([0]), [2])
    with pytest.raises(ValueError):          # wrong dtype
        thg.neighbor_sampling_homogenous(ptrs.int(), idx, dev([0]), [2])
    with pytest.raises(ValueError):
        thg.neighbor_sampling_homogenous(ptrs, idx, dev([0]), [2], thg.WeightedEdgeSampler(torch.ones(idx.numel()).cuda()))
    # the library is still healthy afterwards
    assert thg.neighbor_sampling_homogenous(ptrs, idx, dev([0]), [2])[0].numel() == 3


def test_reservoir_marginals_on_gpu(thg):
    """quirk Q1 closed form, measured on the CUDA path: (k-1)/(n-1) | k/(n-1)."""
    n, k, reps = 16, 5, 40000
    ptrs = np.zeros(n + 2, dtype=np.int64)
    ptrs[1:] = n
    idx = np.arange(1, n + 1)
    thg.rng_reseed(99)
    out = thg.neighbor_sampling_homogenous(dev(ptrs), dev(idx), dev(np.zeros(reps, dtype=np.int64)), [k])
    counts = np.bincount(out[3].cpu().numpy(), minlength=n).astype(np.float64)
    assert counts.sum() == reps * k
    assert chi2_pvalue(counts, reservoir_inclusion(n, k) * reps) > 0.01
    assert chi2_pvalue(counts, np.full(n, k / n) * reps) < 1e-6


def test_gpu_vs_sequential_oracle_histograms(thg, karate):
    """per-frontier-node neighbour histograms vs the reference-order xoshiro oracle, chi-square p > 0.01
    (north_star's statistical bar), for all three samplers."""
    ei, n = karate
    ptrs, idx, _ = graph(thg, ei, n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    reps = 5000
    nodes = [0, 33, 32, 2, 1]
    inputs = np.tile(np.array(nodes), reps)
    w = np.random.default_rng(5).uniform(0.2, 5.0, hi.size)
    cases = [(None, None, 5), (thg.UniformEdgeSampler(True), ("uniform", True), 5),
             (thg.WeightedEdgeSampler(torch.as_tensor(w).cuda()), ("weighted", w), 4)]
    for gs, os_, k in cases:
        thg.rng_reseed(31)
        out = thg.neighbor_sampling_homogenous(ptrs, idx, dev(inputs), [k], gs)
        g = np.bincount(out[3].cpu().numpy(), minlength=hi.size)
        s, r, c, e, lo = O.neighbor_sampling_homogenous(hp, hi, inputs, [k], sampler=os_, rng_mode=O.RNG_XOSHIRO, seed=77)
        o = np.bincount(e, minlength=hi.size)
        for wn in nodes:
            a, b = g[hp[wn]:hp[wn + 1]], o[hp[wn]:hp[wn + 1]]
            assert a.sum() == b.sum()
            assert chi2_two_sample(a, b) > 0.01


def test_calls_are_not_reproducible_unless_reseeded(thg, karate):
    """src/utils/random.rs:8-23: each call forks a fresh child stream."""
    ei, n = karate
    ptrs, idx, _ = graph(thg, ei, n)
    inp = dev(np.arange(34))
    a = thg.neighbor_sampling_homogenous(ptrs, idx, inp, [3, 3])[3]
    b = thg.neighbor_sampling_homogenous(ptrs, idx, inp, [3, 3])[3]
    assert not torch.equal(a, b)
    thg.rng_reseed(4)
    a = thg.neighbor_sampling_homogenous(ptrs, idx, inp, [3, 3])[3]
    thg.rng_reseed(4)
    b = thg.neighbor_sampling_homogenous(ptrs, idx, inp, [3, 3])[3]
    assert torch.equal(a, b)


# ---------------------------------------------------------------------------------------------
# heterogeneous
# ---------------------------------------------------------------------------------------------
def _hetero_graph(thg, fakehetero):
    counts, edges = fakehetero
    node_types = sorted(counts)
    edge_types = sorted(edges)
    cp, ri, hcp, hri = {}, {}, {}, {}
    for et in edge_types:
        k = thg.rel_key(et)
        p, i, _ = thg.to_csc(dev(edges[et]), (counts[et[0]], counts[et[2]]))
        cp[k], ri[k] = p, i
        hcp[k], hri[k] = p.cpu().numpy(), i.cpu().numpy()
        op, oi, _ = O.to_csc(edges[et], (counts[et[0]], counts[et[2]]))
        assert (hcp[k] == op).all() and (hri[k] == oi).all()
    return node_types, edge_types, cp, ri, hcp, hri


def _cmp_hetero(got, want):
    gs, gr, gc, ge, glo = got
    ws, wr, wc, we, wlo = want
    assert set(gs) == set(ws) and set(gr) == set(wr)
    for t in ws:
        assert (gs[t].cpu().numpy() == ws[t]).all(), t
    for k in wr:
        assert (gr[k].cpu().numpy() == wr[k]).all(), k
        assert (gc[k].cpu().numpy() == wc[k]).all(), k
        assert (ge[k].cpu().numpy() == we[k]).all(), k
        assert list(glo[k]) == list(wlo[k]), k


@pytest.mark.parametrize("fan", [[4, 3], [10, 10], [40, 2, 2]])
def test_heterogenous_bit_exact(thg, fakehetero, fan):
    node_types, edge_types, cp, ri, hcp, hri = _hetero_graph(thg, fakehetero)
    inputs = {t: np.array([0, 1, 4, 5]) for t in node_types}
    nn = {thg.rel_key(et): fan for et in edge_types}
    for sampler, osamp in ((None, None), (thg.UniformEdgeSampler(True), ("uniform", True))):
        seed = next_seed(thg, 21)
        got = thg.neighbor_sampling_heterogenous(node_types, edge_types, cp, ri, {t: dev(v) for t, v in inputs.items()},
                                                 nn, len(fan), sampler)
        want = O.neighbor_sampling_heterogenous(node_types, edge_types, hcp, hri, inputs, nn, len(fan), sampler=osamp,
                                                seed=seed)
        _cmp_hetero(got, want)
        for et in edge_types:  # the reference's own invariants, neighbor_sampling.rs:637-646
            k = thg.rel_key(et)
            validate_neighbor_samples(hcp[k], hri[k], want[1][k], want[2][k], want[0][et[0]], want[0][et[2]], want[4][k], fan)


def test_heterogenous_weighted_and_partial(thg, fakehetero):
    node_types, edge_types, cp, ri, hcp, hri = _hetero_graph(thg, fakehetero)
    rng = np.random.default_rng(8)
    w = {k: rng.integers(1, 40, v.size) / 8.0 for k, v in hri.items()}
    # only two relations sampled, seeds for a single node type, different fanouts per relation
    rels = [thg.rel_key(edge_types[0]), thg.rel_key(edge_types[3])]
    nn = {rels[0]: [3, 2], rels[1]: [2, 5]}
    inputs = {edge_types[0][2]: np.array([3, 3, 9, 200])}
    seed = next_seed(thg, 5)
    got = thg.neighbor_sampling_heterogenous(node_types, edge_types, cp, ri, {t: dev(v) for t, v in inputs.items()}, nn, 2,
                                             thg.WeightedEdgeSampler({k: torch.as_tensor(v).cuda() for k, v in w.items()}))
    want = O.neighbor_sampling_heterogenous(node_types, edge_types, hcp, hri, inputs, nn, 2, sampler=("weighted", w), seed=seed)
    _cmp_hetero(got, want)
    for k in got[4]:
        assert (len(got[4][k]) == 2) == (k in nn)


def test_index_replica_does_not_change_results(thg, fakedataset, monkeypatch):
    """the int32 row_indices replica is a memory-layout optimisation only; in-place edits invalidate it."""
    ei, n = fakedataset
    ptrs, idx, _ = graph(thg, ei, n)
    inputs = dev(np.random.default_rng(3).integers(0, n, 300))
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("TCHGEO_INDEX_REPLICA", flag)
        thg.rng_reseed(8)
        outs.append(thg.neighbor_sampling_homogenous(ptrs, idx, inputs, [15, 10, 5]))
    for a, b in zip(outs[0][:4], outs[1][:4]):
        assert torch.equal(a, b)
    # mutate the graph in place: the cached replica must not be used any more
    monkeypatch.setenv("TCHGEO_INDEX_REPLICA", "1")
    idx2 = idx.clone()
    thg.rng_reseed(8)
    a = thg.neighbor_sampling_homogenous(ptrs, idx2, inputs, [4])
    idx2.add_(0).copy_(torch.flip(idx, [0]))
    thg.rng_reseed(8)
    b = thg.neighbor_sampling_homogenous(ptrs, idx2, inputs, [4])
    assert torch.equal(a[3], b[3]) and torch.equal(b[0][300:], idx2[b[3]])


def test_heterogenous_batched_plan(thg, fakehetero):
    node_types, edge_types, cp, ri, hcp, hri = _hetero_graph(thg, fakehetero)
    B = 5
    rng = np.random.default_rng(12)
    inputs = {node_types[0]: rng.integers(0, 800, (B, 6)), node_types[2]: rng.integers(0, 800, (B, 3))}
    nn = {thg.rel_key(et): [3, 2] for et in edge_types}
    plan = thg.HeterogenousSampler(node_types, edge_types, cp, ri, B, {t: v.shape[1] for t, v in inputs.items()}, nn, 2)
    plan.sample({t: dev(v) for t, v in inputs.items()}, seed=77, batch_base=3)
    for b in range(B):
        want = O.neighbor_sampling_heterogenous(node_types, edge_types, hcp, hri, {t: v[b] for t, v in inputs.items()},
                                                nn, 2, seed=77, batch=3 + b)
        _cmp_hetero(plan.batch(b), want)


@pytest.mark.parametrize("kind", ["uniform", "replace", "weighted"])
def test_many_batches_hubs_and_long_lookback_chains(thg, kind):
    """37 batches over a power-law graph with a 3000-neighbour hub and zero-degree nodes: tile-major tickets
    interleave the batches, hop 3 spans several tiles per batch (look-back), heavy nodes take the warp-strided
    draw path.  Two of the batches are checked against the oracle bit for bit, all of them structurally."""
    n = 20000
    ei = _power_law_graph(n, 11, 3000)
    ptrs, idx, _ = graph(thg, ei, n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    B, S, fan = 37, 40, [15, 10, 5]
    inputs = np.random.default_rng(5).integers(0, n, (B, S))
    inputs[3, :5] = 0  # the hub, repeatedly
    sampler, osampler = None, None
    if kind == "replace":
        sampler, osampler = thg.UniformEdgeSampler(True), ("uniform", True)
    elif kind == "weighted":
        w = np.random.default_rng(6).uniform(0.2, 5.0, hi.size)
        sampler, osampler = thg.WeightedEdgeSampler(torch.as_tensor(w).cuda()), ("weighted", w)
    res = thg.neighbor_sampling_homogenous_batched(ptrs, idx, dev(inputs), fan, sampler, seed=4321, batch_base=9)
    for b in range(B):
        got = [t.cpu().numpy() for t in res.batch(b)[:4]] + [res.batch(b)[4]]
        validate_tree_identities(hp, hi, inputs[b], *got, fan, replace=(kind == "replace"))
        if b in (3, B - 1):
            want = O.neighbor_sampling_homogenous(hp, hi, inputs[b], fan, sampler=osampler, seed=4321, batch=9 + b)
            assert_same(got, want)


def test_many_tiles_single_batch(thg, fakedataset):
    """one batch whose last hop spans hundreds of tiles: the look-back walks several 32-wide windows"""
    ei, n = fakedataset
    ptrs, idx, _ = graph(thg, ei, n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    inputs = np.random.default_rng(21).integers(0, n, 400)
    for sampler in (None, thg.UniformEdgeSampler(True)):
        got, seed = run_homo(thg, ptrs, idx, inputs, [16, 8, 4], sampler, state=33)
        want = O.neighbor_sampling_homogenous(hp, hi, inputs, [16, 8, 4], sampler=oracle_sampler(sampler), seed=seed)
        assert_same(got, want)


def test_sample_async_on_two_streams_equals_sample(thg, fakedataset):
    """HomogenousSampler.sample_async()/result(): two plans on two streams, step s+1 enqueued before step s is
    collected, give the same bits as the synchronous call; device-side errors surface in result()."""
    ei, n = fakedataset
    ptrs, idx, _ = graph(thg, ei, n)
    B, S, fan = 9, 21, [15, 10, 5]
    rng = np.random.default_rng(77)
    steps = [dev(rng.integers(0, n, (B, S))) for _ in range(5)]
    ref_plan = thg.HomogenousSampler(ptrs, idx, B, S, fan)
    want = []
    for s, inp in enumerate(steps):
        r = ref_plan.sample(inp, seed=500 + s, batch_base=s * B)
        want.append([[t.clone() for t in r.batch(b)[:4]] + [r.batch(b)[4]] for b in range(B)])
    plans = [thg.HomogenousSampler(ptrs, idx, B, S, fan) for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    pending, got = [], {}

    def take():
        s, j = pending.pop(0)
        r = plans[j].result()
        got[s] = [[t.clone() for t in r.batch(b)[:4]] + [r.batch(b)[4]] for b in range(B)]

    for s, inp in enumerate(steps):
        j = s & 1
        if len(pending) == 2:
            take()
        with torch.cuda.stream(streams[j]):
            plans[j].sample_async(inp, seed=500 + s, batch_base=s * B)
        pending.append((s, j))
    while pending:
        take()
    torch.cuda.synchronize()
    for s in range(len(steps)):
        for b in range(B):
            for x, y in zip(got[s][b][:4], want[s][b][:4]):
                assert torch.equal(x, y)
            assert got[s][b][4] == want[s][b][4]
    bad = steps[0].clone()
    bad[2, 3] = n + 9
    with pytest.raises(thg.ReferencePanic):
        plans[0].sample_async(bad, seed=1).result()


def test_short_edge_attributes_are_refused(thg, fakedataset):
    """weights / timestamps shorter than row_indices: the reference panics on EdgeAttr::get (graph.rs:103-120); a
    col_ptrs that runs past row_indices is caught by the kernels (nnz check) instead of reading out of bounds."""
    ei, n = fakedataset
    ptrs, idx, _ = graph(thg, ei, n)
    inputs = dev(np.arange(64))
    w = torch.ones(idx.numel() - 5, dtype=torch.float64, device="cuda")
    with pytest.raises(thg.ReferencePanic):
        thg.neighbor_sampling_homogenous(ptrs, idx, inputs, [3], thg.WeightedEdgeSampler(w))
    ts = torch.zeros(idx.numel() - 1, dtype=torch.int64, device="cuda")
    flt = (thg.TemporalEdgeFilter((0, 2), ts, True, thg.TEMPORAL_SAMPLE_STATIC), torch.zeros(64, dtype=torch.int64, device="cuda"))
    with pytest.raises(thg.ReferencePanic):
        thg.neighbor_sampling_homogenous(ptrs, idx, inputs, [3], None, flt)
    with pytest.raises(thg.ReferencePanic):
        thg.neighbor_sampling_homogenous(ptrs, idx[:-40].contiguous(), dev(np.arange(n)), [3])
    thg.clear_caches()


def test_temporal_filter_with_large_fanout(thg, fakedataset):
    """fanout above 11 k needs more than the default 48 KB of dynamic shared memory in hop_filtered_kernel"""
    ei, n = fakedataset
    ptrs, idx, _ = graph(thg, ei, n)
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    ts = np.random.default_rng(0).integers(0, 4, hi.size)
    seeds = np.arange(20)
    flt = (thg.TemporalEdgeFilter((0, 2), dev(ts), True, thg.TEMPORAL_SAMPLE_STATIC), dev(np.zeros(20)))
    got = thg.neighbor_sampling_homogenous(ptrs, idx, dev(seeds), [20000], None, flt)
    keep = [(ts[hp[w]:hp[w + 1]] <= 2).sum() for w in seeds]
    assert got[1].numel() == int(np.sum(keep))
