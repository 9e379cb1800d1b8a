"""GPU parity: to_csc / to_csr / ind2ptr vs the reference's KATs and the CPU oracle (bit-exact)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def thg():
    import tch_geometric
    return tch_geometric


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.int64).cuda()


def test_ind2ptr_kat(thg):
    # src/data/storage.rs:153-163
    assert thg.ind2ptr(dev([3, 3, 3, 4, 4, 7, 7, 8, 8]), 10).tolist() == [0, 0, 0, 0, 3, 5, 5, 5, 7, 9, 9]
    assert thg.ind2ptr(dev([]), 4).tolist() == [0, 0, 0, 0, 0]
    assert thg.ind2ptr(dev([0, 0, 9]), 9).tolist() == [0, 2, 2, 2, 2, 2, 2, 2, 2, 2]
    x = np.sort(np.random.default_rng(0).integers(0, 5000, 100000))
    assert (thg.ind2ptr(dev(x), 5000).cpu().numpy() == O.ind2ptr(x, 5000)).all()


def test_to_csc_kat(thg):
    # src/data/storage.rs:166-184
    ei = dev([[1, 2, 3, 4, 9, 5, 6, 7], [0, 0, 0, 1, 4, 1, 2, 2]])
    ptrs, idx, perm = (t.cpu().numpy() for t in thg.to_csc(ei, 10))
    deg = np.diff(ptrs)
    assert (deg[0], deg[1], deg[4], deg[2]) == (3, 2, 1, 2)
    assert idx[ptrs[0]:ptrs[1]].tolist() == [1, 2, 3] and idx[ptrs[1]:ptrs[2]].tolist() == [4, 5]
    assert perm.tolist() == [0, 1, 2, 3, 5, 6, 7, 4]


@pytest.mark.parametrize("csc", [True, False])
def test_fixtures_bit_exact(thg, karate, fakedataset, fakehetero, csc):
    cases = [(karate[0], karate[1]), (fakedataset[0], fakedataset[1])]
    counts, edges = fakehetero
    for (s, _, d), ei in edges.items():
        cases.append((ei, (counts[s], counts[d])))
    rng = np.random.default_rng(0)
    cases.append((np.stack([rng.integers(0, 50, 400), rng.integers(0, 7, 400)]), (50, 7)))       # duplicates (Q9: stable)
    cases.append((np.zeros((2, 0), dtype=np.int64), (5, 9)))                                      # empty
    cases.append((np.array([[0], [0]]), 1))                                                       # 1 node
    cases.append((np.stack([rng.integers(0, 3, 1000), rng.integers(0, 100000, 1000)]), (3, 100000)))  # mostly empty columns
    e = 1_000_000
    cases.append((np.stack([rng.integers(0, 200_000, e), rng.integers(0, 200_000, e)]), 200_000))    # 1M-edge random graph
    for ei, size in cases:
        got = (thg.to_csc if csc else thg.to_csr)(dev(ei), size)
        want = (O.to_csc if csc else O.to_csr)(ei, size)
        for name, g, w in zip(("ptrs", "indices", "perm"), got, want):
            assert (g.cpu().numpy() == w).all(), (name, size)


@pytest.mark.parametrize("csc", [True, False])
def test_partition_form_equals_radix_form(thg, monkeypatch, csc):
    """tchgeo_coo_to_csx has two forms (csx_build.cu): buckets of major ids finished in shared memory, and a device-wide
    radix sort.  Same ptrs / indices / perm, also where the first hands over to the second (a column too long for a
    bucket) and where a whole CTA sorts one column (> 1024 entries)."""
    rng = np.random.default_rng(5)
    fn = thg.to_csc if csc else thg.to_csr
    flip = (lambda a: a) if csc else (lambda a: a[::-1])
    hub = np.stack([rng.integers(0, 50_000, 40_000), np.zeros(40_000, dtype=np.int64)])
    hub[1, :3000] = rng.integers(0, 1000, 3000)                                  # one column of 37 000: radix form takes over
    heavy = np.stack([rng.integers(0, 50_000, 60_000), rng.integers(0, 6, 60_000) * 100])   # six columns of ~10 000
    dup = np.stack([rng.integers(0, 40, 200_000), rng.integers(0, 3000, 200_000)])          # many duplicate edges (Q9)
    skew = np.stack([rng.integers(0, 1_000_000, 300_000), (rng.pareto(1.2, 300_000) * 50).astype(np.int64) % 400_000])
    for ei, size in ((hub, (50_000, 1000)), (heavy, (50_000, 600)), (dup, (40, 3000)), (skew, (1_000_000, 400_000))):
        ei, size = np.ascontiguousarray(flip(ei)), flip(size)
        monkeypatch.setenv("TCHGEO_CSX_SORT", "cub")
        want = fn(dev(ei), size)
        monkeypatch.setenv("TCHGEO_CSX_SORT", "partition")
        got = fn(dev(ei), size)
        for name, g, w in zip(("ptrs", "indices", "perm"), got, want):
            assert torch.equal(g, w), (name, size)
        o = (O.to_csc if csc else O.to_csr)(ei, size)
        assert all((g.cpu().numpy() == w).all() for g, w in zip(got, o))


def test_karate_anchor(thg, karate):
    ei, n = karate
    ptrs, idx, perm = thg.to_csc(dev(ei), n)
    assert ptrs.tolist()[:6] == [0, 16, 25, 35, 41, 44] and ptrs[-1].item() == 156
    assert idx[:16].tolist() == [1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13, 17, 19, 21, 31]
    assert torch.equal(dev(ei)[0][perm], idx)


def test_errors(thg):
    with pytest.raises(thg.ReferencePanic):  # endpoint out of range: the reference's ind2ptr writes out of bounds
        thg.to_csc(dev([[0, 5], [0, 1]]), 3)
    with pytest.raises(ValueError):
        thg.to_csc(torch.zeros((2, 3), dtype=torch.int64), 3)          # CPU tensor
    with pytest.raises(ValueError):
        thg.to_csc(torch.zeros((2, 3), dtype=torch.int32).cuda(), 3)   # wrong dtype
    with pytest.raises(ValueError):
        thg.to_csc(torch.zeros((3, 3), dtype=torch.int64).cuda(), 3)   # wrong shape
