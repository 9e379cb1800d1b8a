"""BASELINE.json full-size configuration (products-shaped: 2 449 029 nodes, 61 859 140 edges, fanouts
[15,10,5], 1024 seeds x many batches) checked through size-independent properties on the device, plus
an exact comparison of one whole batch against the oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tools import synth

pytestmark = pytest.mark.gpu

FAN = [15, 10, 5]


@pytest.fixture(scope="module")
def thg():
    import tch_geometric
    return tch_geometric


@pytest.fixture(scope="module")
def products(thg):
    dev = torch.device("cuda", 0)
    ei, n = synth.products_like(dev)
    ptrs, idx, perm = thg.to_csc(ei, n)
    return ei, n, ptrs, idx, perm


def test_to_csc_properties_at_full_size(products):
    ei, n, ptrs, idx, perm = products
    E = ei.shape[1]
    assert E == synth.PRODUCTS["num_edges"] and n == synth.PRODUCTS["num_nodes"]
    assert ptrs[0].item() == 0 and ptrs[-1].item() == E
    deg = ptrs[1:] - ptrs[:-1]
    assert (deg >= 0).all()
    assert torch.equal(torch.bincount(ei[1], minlength=n), deg)                  # ptrs = histogram of dst
    # perm is a permutation (checksums + range) and reorders the COO input into the CSC arrays
    assert perm.min().item() == 0 and perm.max().item() == E - 1
    assert perm.sum().item() == E * (E - 1) // 2
    assert torch.equal(ei[0][perm], idx)
    col_sorted = ei[1][perm]
    assert (col_sorted[1:] >= col_sorted[:-1]).all()                             # sorted by column
    same_col = col_sorted[1:] == col_sorted[:-1]
    assert (idx[1:][same_col] > idx[:-1][same_col]).all()                        # strictly ascending inside a column
    # idempotence: rebuilding COO from the CSC and converting again is a fixed point
    col_of = torch.repeat_interleave(torch.arange(n, device=ptrs.device), deg)
    import tch_geometric as thg
    p2, i2, perm2 = thg.to_csc(torch.stack([idx, col_of]), n)
    assert torch.equal(p2, ptrs) and torch.equal(i2, idx)
    assert torch.equal(perm2, torch.arange(E, device=ptrs.device))


@pytest.mark.parametrize("sampler_kind", ["uniform", "replace"])
def test_sampling_properties_at_full_size(thg, products, sampler_kind):
    ei, n, ptrs, idx, perm = products
    B, S = 64, 1024
    seeds = torch.from_numpy(synth.seed_batches(n, B, S)).cuda()
    sampler = thg.UniformEdgeSampler(sampler_kind == "replace")
    res = thg.neighbor_sampling_homogenous_batched(ptrs, idx, seeds, FAN, sampler, seed=99)
    deg = ptrs[1:] - ptrs[:-1]
    for b in (0, 17, B - 1):
        samples, rows, cols, eidx, lo = res.batch(b)
        ns, ne = samples.numel(), rows.numel()
        assert ns == S + ne
        assert torch.equal(samples[:S], seeds[b])
        assert torch.equal(rows, torch.arange(S, S + ne, device=rows.device))
        assert torch.equal(samples[rows], idx[eidx])                               # sampled node = CSC entry
        w = samples[cols]
        assert ((ptrs[w] <= eidx) & (eidx < ptrs[w + 1])).all()                      # entry lies in the parent's column
        assert (cols[1:] >= cols[:-1]).all()
        begin, end, e0 = 0, S, 0
        for h, k in enumerate(FAN):
            assert lo[h] == (end, e0, end)
            d = deg[samples[begin:end]]
            expect = torch.where(d > 0, torch.full_like(d, k), torch.zeros_like(d)) if sampler_kind == "replace" \
                else torch.clamp(d, max=k)
            n_e = int(expect.sum().item())
            c = cols[e0:e0 + n_e]
            assert torch.equal(torch.bincount(c - begin, minlength=end - begin), expect)
            if sampler_kind == "uniform":                                          # picks are distinct
                key = c * (idx.numel() + 1) + eidx[e0:e0 + n_e]
                assert torch.unique(key).numel() == n_e
            e0 += n_e
            begin, end = end, end + n_e
        assert e0 == ne
    # different batches draw different samples from the same seeds
    same = thg.neighbor_sampling_homogenous_batched(ptrs, idx, seeds[:1].repeat(2, 1), FAN, sampler, seed=99)
    assert not torch.equal(same.batch(0)[3], same.batch(1)[3])


def test_one_full_batch_bit_exact_vs_oracle(thg, products):
    ei, n, ptrs, idx, perm = products
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    seeds = synth.seed_batches(n, 2, 1024)
    res = thg.neighbor_sampling_homogenous_batched(ptrs, idx, torch.from_numpy(seeds).cuda(), FAN, seed=4242, batch_base=7)
    for b in range(2):
        got = res.batch(b)
        want = O.neighbor_sampling_homogenous(hp, hi, seeds[b], FAN, seed=4242, batch=7 + b)
        for g, w in zip(got[:4], want[:4]):
            assert (g.cpu().numpy() == w).all()
        assert list(got[4]) == list(want[4])


def test_walks_at_scale(thg, products):
    ei, n, ptrs, idx, perm = products
    rp, ci, _ = thg.to_csr(ei, n)
    S, L = 1_000_000, 80
    start = torch.arange(S, device="cuda") % n
    walks, attempts = thg.random_walk(rp, ci, start, L, 1.0, 0.5, seed=1, return_attempts=True)
    assert walks.shape == (S, L + 1) and torch.equal(walks[:, 0], start)
    alive = walks >= 0
    assert (alive[:, :-1] | ~alive[:, 1:]).all()                                    # -1 padding is a suffix
    a, b = walks[:, :-1][alive[:, 1:]], walks[:, 1:][alive[:, 1:]]
    # every step is an edge: binary search b inside row a
    lo, hi = rp[a], rp[a + 1]
    pos = torch.searchsorted(ci, b + a * 0, right=False) if False else None
    sub = torch.randint(0, a.numel(), (200_000,), device="cuda")
    aa, bb = a[sub].cpu().numpy(), b[sub].cpu().numpy()
    hrp, hci = rp.cpu().numpy(), ci.cpu().numpy()
    for x, y in zip(aa[:5000], bb[:5000]):
        s, e = hrp[x], hrp[x + 1]
        i = s + np.searchsorted(hci[s:e], y)
        assert i < e and hci[i] == y
    steps = int(alive[:, 1:].sum().item())
    assert attempts >= steps
    # a 4096-walker slice matches the oracle bit for bit
    o = O.random_walk(hrp, hci, start[:4096].cpu().numpy(), L, 1.0, 0.5, seed=1)
    assert (walks[:4096].cpu().numpy() == o).all()


def test_mag_shaped_hetero_batch_bit_exact_vs_oracle(thg):
    """BASELINE configs[3] at its full shape (ogbn-mag-shaped: paper 736 389 / author 1 134 649 / institution 8 740 /
    field_of_study 59 965 nodes, 4 relations, 21 M edges), 1024 paper seeds per batch, fanouts [10, 10] per relation:
    two batches of the batched plan equal the oracle bit for bit (samples per node type, rows / cols / edge_index and
    layer offsets per relation), and the relabel stage per node type equals the serial map."""
    dev0 = torch.device("cuda", 0)
    counts, edges = synth.mag_like(dev0)
    node_types, edge_types = list(counts), list(edges)
    cp, ri, hcp, hri = {}, {}, {}, {}
    for et, ei in edges.items():
        k = thg.rel_key(et)
        cp[k], ri[k], _ = thg.to_csc(ei, (counts[et[0]], counts[et[2]]))
        hcp[k], hri[k] = cp[k].cpu().numpy(), ri[k].cpu().numpy()
    B, S, H = 3, 1024, 2
    nn = {thg.rel_key(et): [10, 10] for et in edge_types}
    rng = np.random.default_rng(77)
    seeds = np.stack([rng.choice(counts["paper"], S, replace=False) for _ in range(B)])
    plan = thg.HeterogenousSampler(node_types, edge_types, cp, ri, B, {"paper": S}, nn, H, relabel=True)
    plan.sample({"paper": torch.from_numpy(seeds).to(dev0)}, seed=2024, batch_base=5)
    for b in (0, B - 1):
        want = O.neighbor_sampling_heterogenous(node_types, edge_types, hcp, hri, {"paper": seeds[b]}, nn, H,
                                                seed=2024, batch=5 + b)
        got = plan.batch(b)
        for t in node_types:
            assert (got[0][t].cpu().numpy() == want[0][t]).all(), t
        for k in nn:
            for i in (1, 2, 3):
                assert (got[i][k].cpu().numpy() == want[i][k]).all(), (k, i)
            assert [tuple(x) for x in got[4][k]] == [tuple(x) for x in want[4][k]], k
        assert sum(v.numel() for v in got[1].values()) > 50_000          # a real 2-hop sample, not a degenerate one
        for t, (nodes, local) in plan.relabeled(b).items():
            s = got[0][t].cpu().numpy()
            wn, wl = O.unique_relabel(s, S if t == "paper" else 0)
            assert (nodes.cpu().numpy() == wn).all() and (local.cpu().numpy() == wl).all(), t
    thg.clear_caches()
