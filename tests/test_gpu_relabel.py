"""GPU parity: dedup + insertion-order relabel stage vs the oracle (bit-exact)."""
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


def test_small_cases(thg):
    nodes, local = thg.unique_relabel(dev([5, 3, 5, 7, 3, 9, 7, 5]), 3)
    assert nodes.tolist() == [5, 3, 5, 7, 9] and local.tolist() == [2, 1, 2, 3, 1, 4, 3, 2]
    nodes, local = thg.unique_relabel(dev([]), 0)
    assert nodes.numel() == 0 and local.numel() == 0
    nodes, local = thg.unique_relabel(dev([4, 4, 4]), 0)
    assert nodes.tolist() == [4] and local.tolist() == [0, 0, 0]


@pytest.mark.parametrize("fan", [[5, 5], [15, 10, 5]])
def test_sampled_trees(thg, fakedataset, fan):
    ei, n = fakedataset
    ptrs, idx, _ = thg.to_csc(dev(ei), n)
    for seeds in (np.arange(64), np.array([7, 7, 2, 7]), np.random.default_rng(0).integers(0, n, 500)):
        samples, rows, cols, eidx, lo = thg.neighbor_sampling_homogenous(ptrs, idx, dev(seeds), fan)
        nodes, local = thg.unique_relabel(samples, len(seeds))
        wn, wl = O.unique_relabel(samples.cpu().numpy(), len(seeds))
        assert (nodes.cpu().numpy() == wn).all() and (local.cpu().numpy() == wl).all()
        # relabeled edges reference the deduplicated node list consistently
        assert torch.equal(nodes[local], samples)
        assert torch.equal(nodes[local[rows]], idx[eidx])


def _check_tree(samples, n, num_seeds, nodes, local, nodes_len):
    wn, wl = O.unique_relabel(samples[:n], num_seeds)
    assert nodes_len == wn.size
    assert (nodes[:nodes_len] == wn).all() and (local[:n] == wl).all()


@pytest.mark.parametrize("key32", [True, False])
def test_batched_stage_ragged_trees(thg, key32):
    """tchgeo_unique_relabel_batched on padded rows of different lengths: empty, seeds only, duplicated seeds,
    one tile, several tiles -- each tree bit-exact against the serial oracle."""
    rng = np.random.default_rng(5)
    S, stride = 8, 5000
    lens = np.array([8, 8, 9, 1024, 1025, 4999, 5000, 2048, 8, 3000], dtype=np.int64)
    samples = rng.integers(0, 700, (lens.size, stride))          # few distinct ids: long duplicate chains
    samples[1, :8] = [3, 3, 9, 3, 9, 1, 1, 3]                    # duplicated seeds keep all copies, map to the last one
    samples[5, :8] = 42
    if not key32:
        samples[6] += 1 << 40                                     # ids beyond 32 bits need the 64-bit slots
    nodes, local, nlen = thg.unique_relabel_batched(dev(samples), dev(lens), S, key32=key32)
    nodes, local, nlen = nodes.cpu().numpy(), local.cpu().numpy(), nlen.cpu().numpy()
    for b in range(lens.size):
        _check_tree(samples[b], int(lens[b]), S, nodes[b], local[b], int(nlen[b]))


@pytest.mark.parametrize("bound", [700, 701, 5000, 9000, 70_000, 20_000_000])
def test_batched_stage_with_id_bound(thg, bound):
    """A stated id bound selects the direct-address buckets (bucket = id mod NB, slot = id / NB): same answers, for
    bounds that need one bucket, several, and more buckets than the tree size alone would ask for."""
    rng = np.random.default_rng(bound)
    S, stride = 8, 5000
    lens = np.array([0, 8, 9, 1024, 1025, 4999, 5000, 2048, 8, 3000], dtype=np.int64)
    samples = rng.integers(0, min(bound, 700), (lens.size, stride))
    samples[3] = rng.integers(0, bound, stride)                  # the whole id range, high slots included
    samples[4, ::7] = bound - 1
    samples[1, :8] = [3, 3, 9, 3, 9, 1, 1, 3]
    nodes, local, nlen = thg.unique_relabel_batched(dev(samples), dev(lens), S, id_bound=bound)
    nodes, local, nlen = nodes.cpu().numpy(), local.cpu().numpy(), nlen.cpu().numpy()
    for b in range(lens.size):
        _check_tree(samples[b], int(lens[b]), S, nodes[b], local[b], int(nlen[b]))


def test_batched_stage_id_bound_modes_on_large_trees(thg):
    """Trees of > 2^19 positions: with a small bound a pair still packs into one word, with a bound of 2^25 it does not
    (13 slot bits + 20 position bits) and the unpacked direct form runs; both match the hashed form bit for bit."""
    rng = np.random.default_rng(77)
    S, stride = 64, (1 << 19) + 4097
    lens = np.array([stride, stride - 4321], dtype=np.int64)
    samples = rng.integers(0, 2_000_000, (2, stride))
    samples[0, S:S + 1000] = samples[0, :1000][::-1]             # seeds reappear later (and duplicated seeds exist)
    want = thg.unique_relabel_batched(dev(samples), dev(lens), S, key32=True)
    for bound in (2_000_000, 1 << 25):
        got = thg.unique_relabel_batched(dev(samples), dev(lens), S, id_bound=bound)
        assert torch.equal(got[2], want[2])
        for b in range(2):
            k, m = int(want[2][b]), int(lens[b])
            assert torch.equal(got[0][b, :k], want[0][b, :k]) and torch.equal(got[1][b, :m], want[1][b, :m])
    nodes, local, nlen = (x.cpu().numpy() for x in want)
    _check_tree(samples[1], int(lens[1]), S, nodes[1], local[1], int(nlen[1]))


def test_batched_stage_rejects_ids_beyond_the_stated_bound(thg):
    samples = np.arange(64, dtype=np.int64).reshape(2, 32)
    with pytest.raises(thg.ReferencePanic):
        thg.unique_relabel_batched(dev(samples), dev([32, 32]), 4, id_bound=63)
    samples[1, 5] = -3
    with pytest.raises(thg.ReferencePanic):
        thg.unique_relabel_batched(dev(samples), dev([32, 32]), 4, id_bound=64)
    with pytest.raises(ValueError):
        thg.unique_relabel_batched(dev(samples), dev([32, 32]), 4, id_bound=1 << 33)


def test_batched_stage_rejects_ids_beyond_32_bits_in_key32_mode(thg):
    samples = np.arange(64, dtype=np.int64).reshape(2, 32)
    samples[1, 5] = (1 << 32) + 7
    with pytest.raises(thg.ReferencePanic):
        thg.unique_relabel_batched(dev(samples), dev([32, 32]), 4, key32=True)


def test_sampler_with_relabel_products_size(thg):
    """The stage inside the plan (HomogenousSampler(relabel=True)) on products-size trees (~0.6 M ids per batch, 1024
    seeds, several L2 waves): reference-layout outputs unchanged, nodes / local bit-exact against the oracle."""
    from tools import synth
    ei, n = synth.products_like(torch.device("cuda", 0))
    ptrs, idx, _ = thg.to_csc(ei, n)
    del ei
    B, S, fan = 12, 1024, [15, 10, 5]
    seeds = synth.seed_batches(n, B, S)
    seeds[3, 100:200] = seeds[3, :100]                            # duplicated seeds
    plain = thg.HomogenousSampler(ptrs, idx, B, S, fan).sample(dev(seeds), seed=11)
    plan = thg.HomogenousSampler(ptrs, idx, B, S, fan, relabel=True)
    res = plan.sample(dev(seeds), seed=11)
    assert (res.samples_len == plain.samples_len).all() and (res.edges_len == plain.edges_len).all()
    for b in (0, 3, B - 1):
        for g, w in zip(res.batch(b)[:4], plain.batch(b)[:4]):
            assert torch.equal(g, w)
        ns = int(res.samples_len[b])
        nodes, local = res.relabeled(b)
        _check_tree(res.samples[b].cpu().numpy(), ns, S, nodes.cpu().numpy(), local.cpu().numpy(), nodes.numel())
        samples, rows, cols, eidx, _ = res.batch(b)
        assert torch.equal(nodes[local], samples)
        assert torch.equal(nodes[local[rows]], idx[eidx])         # relabelled edge endpoints are the sampled CSC entries
        assert nodes.numel() == torch.unique(samples[S:]).numel() + S - int(
            torch.isin(torch.unique(samples[S:]), samples[:S]).sum().item())
    # the async pair gives the same result
    plan.sample_async(dev(seeds), seed=11)
    again = plan.result()
    assert (again.nodes_len == res.nodes_len).all()
    thg.clear_caches()


def test_hetero_sampler_with_relabel(thg, fakehetero):
    """K7 per node type: the batched heterogeneous plan relabels every type's samples vector."""
    counts, edges = fakehetero
    node_types, edge_types = list(counts), list(edges)
    cp, ri = {}, {}
    for et, e in edges.items():
        cp[thg.rel_key(et)], ri[thg.rel_key(et)], _ = thg.to_csc(dev(e), (counts[et[0]], counts[et[2]]))
    B, S = 3, 16
    rng = np.random.default_rng(2)
    inputs = {t: dev(rng.integers(0, counts[t], (B, S))) for t in node_types}
    nn = {thg.rel_key(et): [4, 3] for et in edge_types}
    plan = thg.HeterogenousSampler(node_types, edge_types, cp, ri, B, {t: S for t in node_types}, nn, 2, relabel=True)
    plan.sample(inputs, seed=9)
    for b in range(B):
        out_s = plan.batch(b)[0]
        rl = plan.relabeled(b)
        for t in node_types:
            nodes, local = rl[t]
            s = out_s[t].cpu().numpy()
            _check_tree(s, s.size, S, nodes.cpu().numpy(), local.cpu().numpy(), nodes.numel())
