"""Shared checkers; validate_neighbor_samples restates the reference's test helper
(src/algo/neighbor_sampling.rs:370-401) and extends it with the structural identities of SURVEY §8(c)."""
import numpy as np


def has_edge(ptrs, indices, x, y):
    lo, hi = int(ptrs[x]), int(ptrs[x + 1])
    i = lo + int(np.searchsorted(indices[lo:hi], y))
    return i < hi and indices[i] == y


def validate_neighbor_samples(ptrs, indices, rows, cols, samples_src, samples_dst, layer_offsets, num_neighbors):
    """neighbor_sampling.rs:370-401"""
    for j, i in zip(rows, cols):
        v, w = samples_src[j], samples_dst[i]
        assert has_edge(ptrs, indices, w, v)  # csc: dst <- src
    counts = np.zeros(len(samples_dst), dtype=np.int64)
    np.add.at(counts, cols, 1)
    begin = 0
    for h, (_, _, dst_end) in enumerate(layer_offsets):
        assert (counts[begin:dst_end] <= num_neighbors[h]).all()
        begin = dst_end


def validate_tree_identities(ptrs, indices, inputs, samples, rows, cols, edge_index, layer_offsets, num_neighbors,
                             replace=False):
    """Homogeneous structural identities: rows is an arange, edge_index are CSC positions
    inside the parent's column, per-node counts are min(deg,k) / k*[deg>0], picks distinct w/o replacement."""
    S, E = len(inputs), len(rows)
    assert len(samples) == S + E
    assert (samples[:S] == inputs).all()
    assert (rows == np.arange(S, S + E)).all()
    assert (samples[rows] == indices[edge_index]).all()
    w = samples[cols]
    assert (ptrs[w] <= edge_index).all() and (edge_index < ptrs[w + 1]).all()
    assert (np.diff(cols) >= 0).all()  # frontier order is preserved
    deg = ptrs[1:] - ptrs[:-1]
    begin, end = 0, S
    e0 = 0
    for h, k in enumerate(num_neighbors):
        assert layer_offsets[h] == (end, e0, end)
        front = samples[begin:end]
        d = deg[front]
        expect = np.where(d > 0, k, 0) if replace else np.minimum(d, k)
        n_e = int(expect.sum())
        c = cols[e0:e0 + n_e]
        assert (np.bincount(c - begin, minlength=end - begin) == expect).all() if n_e else True
        if not replace:
            key = c * (len(indices) + 1) + edge_index[e0:e0 + n_e]
            assert len(np.unique(key)) == n_e
        e0 += n_e
        begin, end = end, end + n_e
    assert e0 == E


def full_neighborhood_tree(ptrs, indices, inputs, num_hops):
    """The single exact answer of fanout >= max degree without replacement (deterministic regime)."""
    samples = list(int(x) for x in inputs)
    rows, cols, eidx, lo = [], [], [], []
    begin, end = 0, len(samples)
    for _ in range(num_hops):
        lo.append((len(samples), len(cols), len(samples)))
        for i in range(begin, end):
            w = samples[i]
            for p in range(int(ptrs[w]), int(ptrs[w + 1])):
                rows.append(len(samples))
                samples.append(int(indices[p]))
                cols.append(i)
                eidx.append(p)
        begin, end = end, len(samples)
    a = lambda x: np.asarray(x, dtype=np.int64)
    return a(samples), a(rows), a(cols), a(eidx), lo


def chi2_pvalue(observed, expected):
    from scipy import stats
    observed = np.asarray(observed, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    m = expected > 0
    assert (observed[~m] == 0).all()
    stat = ((observed[m] - expected[m]) ** 2 / expected[m]).sum()
    return float(stats.chi2.sf(stat, int(m.sum()) - 1))


def chi2_two_sample(a, b):
    """Chi-square homogeneity test of two count vectors."""
    from scipy import stats
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m = (a + b) > 0
    if m.sum() < 2:
        return 1.0
    _, p, _, _ = stats.chi2_contingency(np.stack([a[m], b[m]]))
    return float(p)


def reservoir_inclusion(n, k):
    """Closed-form marginals of the reference's off-by-one reservoir (quirk Q1, sampling.rs:17-19):
    (k-1)/(n-1) for the first k items, k/(n-1) for the rest (n > k)."""
    p = np.full(n, k / (n - 1.0))
    p[:k] = (k - 1.0) / (n - 1.0)
    return p
