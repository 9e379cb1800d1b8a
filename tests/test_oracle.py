"""Pins the CPU oracle (oracle/tchgeo_oracle.c) against the reference's own known-answer tests,
fixtures and invariants.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from helpers import (chi2_pvalue, chi2_two_sample, full_neighborhood_tree, has_edge, reservoir_inclusion,
                     validate_neighbor_samples, validate_tree_identities)

KARATE_COLPTR = [0, 16, 25, 35, 41, 44, 48, 52, 56, 61, 63, 66, 67, 69, 74, 76, 78, 80, 82, 84, 87, 89, 91, 93, 98,
                 101, 104, 106, 110, 113, 117, 121, 127, 139, 156]


# ---------------------------------------------------------------------------------------------
# RNG primitives
# ---------------------------------------------------------------------------------------------
def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_xoshiro256pp_known_answers():
    # the vector published with the reference C implementation (xoshiro256plusplus.c, state {1,2,3,4}); rand 0.8.5's
    # SmallRng carries the same vector as its own unit test (rand is un-vendored here: Cargo.lock:619)
    assert O.kat_xoshiro([1, 2, 3, 4], 10) == [
        41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205, 9973669472204895162,
        14011001112246962877, 12406186145184390807, 15849039046786891736, 10450023813501588000]


def test_seed_from_u64_is_splitmix64():
    # SplitMix64's published outputs: seed 0 starts 0xE220A8397B1DCDAF; seed 1234567 is the vector of the reference C file
    assert O.kat_seed_from_u64(0)[:2] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4]
    assert O.kat_seed_from_u64(1234567) == [6457827717110365317, 3203168211198807973, 9817491932198370423,
                                            4593380528125082431]


class _PyXoshiro:
    """independent restatement in Python integers of rand 0.8.5's reductions"""
    M = (1 << 64) - 1

    def __init__(self, seed):
        self.s = O.kat_seed_from_u64(seed)

    def next_u64(self):
        s, M = self.s, self.M
        rotl = lambda x, k: ((x << k) | (x >> (64 - k))) & M
        out = (rotl((s[0] + s[3]) & M, 23) + s[0]) & M
        t = (s[1] << 17) & M
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45)
        return out

    def gen_range(self, rng):       # UniformInt::sample_single: widening multiply, rejection zone
        zone = ((rng << (64 - rng.bit_length())) & self.M) - 1
        while True:
            m = self.next_u64() * rng
            if (m & self.M) <= zone:
                return m >> 64

    def gen_f32(self):              # upper 32 bits, 23 mantissa bits
        while True:
            v = np.uint32((self.next_u64() >> 32 >> 9) | 0x3F800000).view(np.float32)
            r = np.float32(v - np.float32(1.0))
            if r < 1.0:
                return int(r.view(np.uint32))

    def gen_f64(self, high):        # 52 mantissa bits
        while True:
            v = np.uint64((self.next_u64() >> 12) | 0x3FF0000000000000).view(np.float64)
            r = (v - 1.0) * high
            if r < high:
                return int(np.float64(r).view(np.uint64))


@pytest.mark.parametrize("rng", [1, 2, 3, 5, 17, 1000, (1 << 33) + 7, (1 << 63) + 12345, 3 << 62])
def test_gen_range_restatement(rng):
    py = _PyXoshiro(7)
    got = O.kat_reduce(7, 0, 200, range_=rng)
    assert got == [py.gen_range(rng) for _ in range(200)] and max(got) < rng


def test_gen_float_restatement():
    py = _PyXoshiro(9)
    assert O.kat_reduce(9, 1, 200) == [py.gen_f32() for _ in range(200)]
    py = _PyXoshiro(11)
    assert O.kat_reduce(11, 2, 200, high=3.75) == [py.gen_f64(3.75) for _ in range(200)]


# ---------------------------------------------------------------------------------------------
# CSC/CSR build: reference KATs (src/data/storage.rs:153-184) and fixtures
# ---------------------------------------------------------------------------------------------
def test_ind2ptr_kat():
    assert O.ind2ptr([3, 3, 3, 4, 4, 7, 7, 8, 8], 10).tolist() == [0, 0, 0, 0, 3, 5, 5, 5, 7, 9, 9]


def test_ind2ptr_empty():
    assert O.ind2ptr([], 5).tolist() == [0] * 6


def test_to_csc_kat():
    ei = np.array([[1, 2, 3, 4, 9, 5, 6, 7], [0, 0, 0, 1, 4, 1, 2, 2]])
    ptrs, idx, perm = O.to_csc(ei, 10)
    deg = np.diff(ptrs)
    assert (deg[0], deg[1], deg[4], deg[2]) == (3, 2, 1, 2)
    assert idx[ptrs[0]:ptrs[1]].tolist() == [1, 2, 3]
    assert idx[ptrs[1]:ptrs[2]].tolist() == [4, 5]


def test_karate_anchors(karate):
    ei, n = karate
    ptrs, idx, perm = O.to_csc(ei, n)
    assert ptrs.tolist() == KARATE_COLPTR
    assert idx[:16].tolist() == [1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13, 17, 19, 21, 31]


def _numpy_csx(ei, size, csc):
    s0, s1 = size
    row, col = ei[0], ei[1]
    key = col * s0 + row if csc else row * s1 + col
    perm = np.argsort(key, kind="stable")
    major = (col if csc else row)[perm]
    n_major = s1 if csc else s0
    ptrs = np.searchsorted(major, np.arange(n_major + 1), side="left")
    return ptrs, (row if csc else col)[perm], perm


@pytest.mark.parametrize("csc", [True, False])
def test_to_csx_matches_numpy(karate, fakedataset, fakehetero, csc):
    cases = [(karate[0], (karate[1],) * 2), (fakedataset[0], (fakedataset[1],) * 2)]
    counts, edges = fakehetero
    for (s, _, d), ei in edges.items():
        cases.append((ei, (counts[s], counts[d])))
    rng = np.random.default_rng(0)
    cases.append((np.stack([rng.integers(0, 50, 400), rng.integers(0, 7, 400)]), (50, 7)))  # duplicates, rectangular
    cases.append((np.zeros((2, 0), dtype=np.int64), (5, 9)))                                 # empty
    for ei, size in cases:
        got = (O.to_csc if csc else O.to_csr)(ei, size)
        want = _numpy_csx(ei, size, csc)
        for g, w in zip(got, want):
            assert (g == w).all()


def test_csc_edge_cumsum_kat():
    # src/data/transform.rs:85-97
    got = O.csc_edge_cumsum([0, 0, 0, 0, 3, 5, 5, 5, 7, 9], [9.0, 5.0, 8.0, 9.0, 10.0, 11.0, 1.0, 1.5])
    assert got.tolist() == [9.0, 14.0, 22.0, 9.0, 19.0, 11.0, 12.0, 1.5]


def test_csc_sort_edges_kat():
    # src/data/transform.rs:69-82
    got = O.csc_sort_edges([0, 0, 0, 0, 3, 5, 5, 5, 7, 9], [0, 1, 2, 3, 4, 5, 6, 7],
                           [9.0, 5.0, 8.0, 9.0, 10.0, 11.0, 1.0, 1.5])
    assert got.tolist() == [1, 2, 0, 3, 4, 6, 5, 7]
    desc = O.csc_sort_edges([0, 3, 5], [10, 11, 12, 13, 14], [1.0, 3.0, 2.0, 5.0, 5.0], descending=True)
    assert desc.tolist() == [11, 12, 10, 13, 14]  # ties keep CSC order


# ---------------------------------------------------------------------------------------------
# neighbor sampling: the reference's invariant tests (neighbor_sampling.rs:438-495, :573-648)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def karate_csc(karate):
    ei, n = karate
    return O.to_csc(ei, n)


@pytest.mark.parametrize("mode", [O.RNG_XOSHIRO, O.RNG_COUNTER])
@pytest.mark.parametrize("sampler", [None, ("uniform", True), ("uniform", False), "weighted"])
def test_homogenous_invariants(karate_csc, mode, sampler):
    ptrs, idx, _ = karate_csc
    if sampler == "weighted":  # neighbor_sampling.rs:475: weights ~ U(0.2, 5.0) f64
        sampler = ("weighted", np.random.default_rng(1).uniform(0.2, 5.0, idx.size))
    inputs, fan = np.array([0, 1, 4, 5]), [4, 3]
    for seed in range(5):
        s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, inputs, fan, sampler=sampler, rng_mode=mode, seed=seed)
        validate_neighbor_samples(ptrs, idx, r, c, s, s, lo, fan)
        replace = sampler == ("uniform", True)
        validate_tree_identities(ptrs, idx, inputs, s, r, c, e, lo, fan, replace=replace)


@pytest.mark.parametrize("mode", [O.RNG_XOSHIRO, O.RNG_COUNTER])
def test_deterministic_regime_is_exact(karate_csc, mode):
    """fanout >= max degree without replacement: one exact answer for all five outputs."""
    ptrs, idx, _ = karate_csc
    inputs = np.arange(34)
    want = full_neighborhood_tree(ptrs, idx, inputs, 2)
    for sampler in (None, ("weighted", np.ones(idx.size))):
        got = O.neighbor_sampling_homogenous(ptrs, idx, inputs, [17, 17], sampler=sampler, rng_mode=mode, seed=7)
        for g, w in zip(got[:4], want[:4]):
            assert (g == w).all()
        assert got[4] == want[4] == [(34, 0, 34), (190, 156, 190)]


def test_fanout_zero_panics_like_reference(karate_csc):
    ptrs, idx, _ = karate_csc
    with pytest.raises(O.OraclePanic):  # gen_range(0..0), sampling.rs:19
        O.neighbor_sampling_homogenous(ptrs, idx, [0], [0])
    s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, [0], [0], sampler=("uniform", True))
    assert len(s) == 1 and len(r) == 0
    with pytest.raises(O.OraclePanic):  # out-of-range seed (quirk Q10)
        O.neighbor_sampling_homogenous(ptrs, idx, [34], [2])


def _star(n):
    """column 0 has in-neighbours 1..n"""
    ptrs = np.zeros(n + 2, dtype=np.int64)
    ptrs[1:] = n
    return ptrs, np.arange(1, n + 1, dtype=np.int64)


@pytest.mark.parametrize("mode", [O.RNG_XOSHIRO, O.RNG_COUNTER])
def test_reservoir_bias_matches_closed_form(mode):
    """quirk Q1: inclusion (k-1)/(n-1) for the first k neighbours, k/(n-1) for the rest."""
    n, k, reps = 16, 5, 20000
    ptrs, idx = _star(n)
    counts = np.zeros(n)
    if mode == O.RNG_COUNTER:
        # independent draws come from distinct (pos, batch) counters: many seeds at once
        s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, np.zeros(reps, dtype=np.int64), [k], rng_mode=mode, seed=3)
        np.add.at(counts, e, 1)
    else:
        s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, np.zeros(reps, dtype=np.int64), [k], rng_mode=mode, seed=3)
        np.add.at(counts, e, 1)
    assert counts.sum() == reps * k
    p = reservoir_inclusion(n, k)
    assert abs(counts[:k].sum() / reps / k - (k - 1) / (n - 1)) < 0.01
    assert chi2_pvalue(counts, p * reps) > 0.001
    # and it is NOT the ideal uniform k/n sampler
    assert chi2_pvalue(counts, np.full(n, k / n) * reps) < 1e-6


def test_counter_and_xoshiro_modes_agree_statistically(karate_csc):
    """The Philox counter layout and the sequential xoshiro restatement draw from the same law:
    per-(parent, neighbour) histograms over many repeats (SURVEY §8c iii)."""
    ptrs, idx, _ = karate_csc
    reps = 4000
    inputs = np.tile(np.array([0, 33, 32, 2]), reps)
    for sampler, k in ((None, 5), (("uniform", True), 5), (("weighted", np.random.default_rng(5).uniform(0.2, 5.0, idx.size)), 4)):
        hist = []
        for mode in (O.RNG_XOSHIRO, O.RNG_COUNTER):
            s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, inputs, [k], sampler=sampler, rng_mode=mode, seed=11)
            hist.append(np.bincount(e, minlength=idx.size))
        for w in (0, 33, 32, 2):
            a, b = hist[0][ptrs[w]:ptrs[w + 1]], hist[1][ptrs[w]:ptrs[w + 1]]
            assert a.sum() == b.sum()
            assert chi2_two_sample(a, b) > 0.001


def test_replacement_always_k(karate_csc):
    """quirk Q3: exactly k picks even when deg < k; zero-degree nodes yield nothing."""
    ptrs, idx, _ = karate_csc
    s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, [11], [6], sampler=("uniform", True), seed=2)  # deg(11)=1
    assert len(r) == 6 and (e == ptrs[11]).all()
    ptrs2 = np.array([0, 0, 2], dtype=np.int64)
    s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs2, [0, 0], [0, 1], [3], sampler=("uniform", True))
    assert c.tolist() == [1, 1, 1]


def test_weighted_bias_matches_q2():
    """quirk Q2 (sampling.rs:47-52): unit weights, n=8, k=3 -> first three ~0.74, tail 0.20 -> 0.12."""
    n, k, reps = 8, 3, 30000
    ptrs, idx = _star(n)
    w = np.ones(n)
    for mode in (O.RNG_XOSHIRO, O.RNG_COUNTER):
        s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, np.zeros(reps, dtype=np.int64), [k],
                                                        sampler=("weighted", w), rng_mode=mode, seed=5)
        inc = np.bincount(e, minlength=n) / reps
        # exact marginals by dynamic programming over the serial algorithm
        p_keep = np.ones(n)
        incl = np.zeros(n)
        surv = np.ones(k)  # P(slot still holds its initial item)
        for i in range(k, n):
            fire = 1.0 / (i + 1)
            for j in range(k, i):
                incl[j] *= (1 - fire / k)
            incl[i] = fire
            surv *= (1 - fire / k)
        incl[:k] = surv
        assert np.abs(inc - incl).max() < 0.012
        assert inc[0] > 0.7 and inc[-1] < 0.15


def test_heterogenous_invariants(fakehetero):
    """neighbor_sampling.rs:573-648: inputs [0,1,4,5] per type, fanouts [4,3] per relation."""
    counts, edges = fakehetero
    node_types = sorted(counts)
    edge_types = sorted(edges)
    cp, ri = {}, {}
    for et in edge_types:
        p, i, _ = O.to_csc(edges[et], (counts[et[0]], counts[et[2]]))
        cp[O.rel_key(et)], ri[O.rel_key(et)] = p, i
    inputs = {t: np.array([0, 1, 4, 5]) for t in node_types}
    nn = {O.rel_key(et): [4, 3] for et in edge_types}
    for mode in (O.RNG_XOSHIRO, O.RNG_COUNTER):
        s, r, c, e, lo = O.neighbor_sampling_heterogenous(node_types, edge_types, cp, ri, inputs, nn, 2, rng_mode=mode, seed=4)
        for et in edge_types:
            k = O.rel_key(et)
            validate_neighbor_samples(cp[k], ri[k], r[k], c[k], s[et[0]], s[et[2]], lo[k], nn[k])
            assert (s[et[0]][r[k]] == ri[k][e[k]]).all()
            assert len(lo[k]) == 2
        # every sampled node of a type comes from exactly one edge of a relation with that src type
        for t in node_types:
            n_edges = sum(len(r[O.rel_key(et)]) for et in edge_types if et[0] == t)
            assert len(s[t]) == 4 + n_edges
            allrows = np.concatenate([r[O.rel_key(et)] for et in edge_types if et[0] == t])
            assert sorted(allrows.tolist()) == list(range(4, 4 + n_edges))


def test_heterogenous_partial_relations(fakehetero):
    """relations absent from num_neighbors get empty outputs and no layer offsets (:280-285)."""
    counts, edges = fakehetero
    node_types = sorted(counts)
    edge_types = sorted(edges)
    cp, ri = {}, {}
    for et in edge_types:
        p, i, _ = O.to_csc(edges[et], (counts[et[0]], counts[et[2]]))
        cp[O.rel_key(et)], ri[O.rel_key(et)] = p, i
    only = O.rel_key(edge_types[0])
    s, r, c, e, lo = O.neighbor_sampling_heterogenous(node_types, edge_types, cp, ri, {edge_types[0][2]: np.array([3, 3])},
                                                      {only: [2, 2]}, 2)
    assert set(r) == set(cp)
    for k in cp:
        if k != only:
            assert len(r[k]) == 0 and lo[k] == []
    assert len(lo[only]) == 2


# ---------------------------------------------------------------------------------------------
# random walk (random_walk.rs:302-331)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [O.RNG_XOSHIRO, O.RNG_COUNTER])
def test_random_walk_invariants(karate, mode):
    ei, n = karate
    rp, ci, _ = O.to_csr(ei, n)
    start = np.array([0, 1, 2, 3])
    walks = O.random_walk(rp, ci, start, 10, 1.0, 1.5, rng_mode=mode, seed=0)
    assert walks.shape == (4, 11)
    assert (walks[:, 0] == start).all()
    for w in walks:
        for a, b in zip(w[:-1], w[1:]):
            assert has_edge(rp, ci, a, b)


def test_random_walk_dead_end_pads_minus_one():
    # 0 -> 1 -> 2, node 2 is a sink
    rp, ci, _ = O.to_csr(np.array([[0, 1], [1, 2]]), 3)
    for mode in (O.RNG_XOSHIRO, O.RNG_COUNTER):
        w = O.random_walk(rp, ci, [0, 2], 4, 1.0, 1.0, rng_mode=mode)
        assert w.tolist() == [[0, 1, 2, -1, -1], [2, -1, -1, -1, -1]]


def test_random_walk_modes_agree_statistically(karate):
    """second-order transition histograms (prev, cur) -> next agree between the two RNG modes and
    follow node2vec's p/q law."""
    ei, n = karate
    rp, ci, _ = O.to_csr(ei, n)
    start = np.tile(np.arange(n), 600)
    hists = []
    for mode in (O.RNG_XOSHIRO, O.RNG_COUNTER):
        w = O.random_walk(rp, ci, start, 3, 0.5, 2.0, rng_mode=mode, seed=9)
        key = (w[:, 1] * n + w[:, 2]) * n + w[:, 3]
        hists.append(np.bincount(key, minlength=n ** 3))
    assert chi2_two_sample(hists[0], hists[1]) > 0.001
    # law check on one (prev, cur) pair: unnormalised weights 1/p (back), 1 (dist 1), 1/q (dist 2)
    prev, cur = 0, 1
    nb = ci[rp[cur]:rp[cur + 1]]
    wts = np.array([2.0 if v == prev else (1.0 if has_edge(rp, ci, v, prev) else 0.5) for v in nb])
    obs = np.array([hists[1][(prev * n + cur) * n + v] for v in nb], dtype=np.float64)
    assert chi2_pvalue(obs, wts / wts.sum() * obs.sum()) > 0.001


# ---------------------------------------------------------------------------------------------
# dedup / relabel stage (negative_sampling.rs:20-47 semantic)
# ---------------------------------------------------------------------------------------------
def _relabel_py(samples, num_seeds):
    nodes = list(samples[:num_seeds])
    mapping = {}
    for i, s in enumerate(nodes):
        mapping[s] = i  # HashMap::extend: later duplicates overwrite
    for s in samples[num_seeds:]:
        if s not in mapping:
            mapping[s] = len(nodes)
            nodes.append(s)
    return nodes, [mapping[s] for s in samples]


def test_unique_relabel_matches_hashmap_semantic(karate_csc):
    ptrs, idx, _ = karate_csc
    assert [x.tolist() for x in O.unique_relabel([5, 3, 5, 7, 3, 9, 7, 5], 3)] == [[5, 3, 5, 7, 9], [2, 1, 2, 3, 1, 4, 3, 2]]
    for seeds in ([0, 1, 4, 5], [7, 7, 2], list(range(34))):
        s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, seeds, [5, 5], seed=1)
        nodes, local = O.unique_relabel(s, len(seeds))
        wn, wl = _relabel_py(s.tolist(), len(seeds))
        assert nodes.tolist() == wn and local.tolist() == wl
        assert (nodes[local] == s).all()


def test_cpu_baseline_driver_matches_single_calls(karate_csc):
    ptrs, idx, _ = karate_csc
    inputs = np.arange(32).reshape(8, 4)
    ts, te = O.neighbor_sampling_homogenous_batches(ptrs, idx, inputs, [5, 5], rng_mode=O.RNG_COUNTER, seed=3, num_threads=2)
    ns = ne = 0
    for b in range(8):
        s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, inputs[b], [5, 5], rng_mode=O.RNG_COUNTER, seed=3, batch=b)
        ns += len(s)
        ne += len(r)
    assert (ts, te) == (ns, ne)


# ---------------------------------------------------------------------------------------------
# temporal filter (neighbor_sampling.rs:36-77), the reference's test :498-570
# ---------------------------------------------------------------------------------------------
def _paths(rows, cols, num_inputs):
    """root index and edge list of every sampled node (tree layout: cols[e] is the parent of rows[e])."""
    parent = {int(j): (int(i), e) for e, (j, i) in enumerate(zip(rows, cols))}
    out = {}
    for j in parent:
        edges, cur = [], j
        while cur >= num_inputs:
            cur, e = parent[cur]
            edges.append(e)
        out[j] = (cur, edges)
    return out


@pytest.mark.parametrize("mode", [O.RNG_XOSHIRO, O.RNG_COUNTER])
def test_temporal_filter_windows(karate_csc, mode):
    ptrs, idx, _ = karate_csc
    ts = np.random.default_rng(0).integers(0, 4, idx.size)          # edge timestamps in [0, 4), :504
    inputs, t0, fan = np.array([0, 1, 4, 5]), np.array([0, 1, 2, 3]), [4, 3]
    # static window 0..=2 (:512-536): every edge on every path has a timestamp inside the window
    flt = dict(mode=O.TEMPORAL_STATIC, forward=False, window=(0, 2), timestamps=ts, inputs_state=t0)
    s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, inputs, fan, filter=flt, rng_mode=mode, seed=1)
    validate_neighbor_samples(ptrs, idx, r, c, s, s, lo, fan)
    assert len(e) > 0 and ((ts[e] >= 0) & (ts[e] <= 2)).all()
    # relative backward window (:538-569): start_t - 2 <= t <= start_t along the whole path
    flt = dict(mode=O.TEMPORAL_RELATIVE, forward=False, window=(0, 2), timestamps=ts, inputs_state=t0)
    s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, inputs, fan, filter=flt, rng_mode=mode, seed=2)
    validate_neighbor_samples(ptrs, idx, r, c, s, s, lo, fan)
    for j, (root, edges) in _paths(r, c, len(inputs)).items():
        for ed in edges:
            assert t0[root] - 2 <= ts[e[ed]] <= t0[root]
    # dynamic forward: each hop is relative to the timestamp of the edge that reached the node
    flt = dict(mode=O.TEMPORAL_DYNAMIC, forward=True, window=(0, 1), timestamps=ts, inputs_state=t0)
    s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, inputs, fan, filter=flt, rng_mode=mode, seed=3)
    hop1 = lo[1][1]
    for ed in range(len(e)):
        prev_t = t0[c[ed]] if ed < hop1 else ts[e[c[ed] - len(inputs)]]
        assert 0 <= ts[e[ed]] - prev_t <= 1


def test_temporal_filter_counts_only_passing_edges(karate_csc):
    """fully filtered neighbourhoods yield nothing (quirk Q4); with replacement picks only passing edges."""
    ptrs, idx, _ = karate_csc
    ts = np.arange(idx.size) % 5
    flt = dict(mode=O.TEMPORAL_STATIC, forward=False, window=(9, 9), timestamps=ts, inputs_state=np.zeros(3, dtype=np.int64))
    s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, [0, 1, 2], [3, 3], filter=flt)
    assert len(r) == 0 and lo == [(3, 0, 3), (3, 0, 3)]
    flt["window"] = (2, 2)
    s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, [0, 1, 2], [6], sampler=("uniform", True), filter=flt)
    assert len(e) == 18 and (ts[e] == 2).all()
