"""GPU parity: gather_rows against numpy fancy indexing, bit-exact for every dtype / row width,
including rows that force the 8-, 4- and 1-byte vector paths."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def thg():
    import tch_geometric
    return tch_geometric


@pytest.mark.parametrize("dtype,width", [(np.float32, 100), (np.float32, 128), (np.float32, 3), (np.float32, 1),
                                         (np.float64, 7), (np.int64, 1), (np.int16, 5), (np.uint8, 13), (np.uint8, 1)])
def test_matches_numpy(thg, dtype, width):
    rng = np.random.default_rng(width)
    src = (rng.standard_normal((5000, width)) * 100).astype(dtype)
    idx = rng.integers(0, 5000, 123457)
    got = thg.gather_rows(torch.from_numpy(src).cuda(), torch.from_numpy(idx).cuda())
    assert got.dtype == torch.from_numpy(src).dtype and tuple(got.shape) == (idx.size, width)
    assert (got.cpu().numpy() == src[idx]).all()


def test_shapes_views_and_errors(thg):
    src = torch.arange(24 * 6, dtype=torch.float32, device="cuda").reshape(24, 2, 3)
    idx = torch.tensor([5, 0, 23, 5], device="cuda")
    assert torch.equal(thg.gather_rows(src, idx), src[idx])
    vec = torch.arange(10, dtype=torch.int64, device="cuda") * 3   # perm[edge_index]
    assert torch.equal(thg.gather_rows(vec, idx[:2]), vec[idx[:2]])
    off = src.reshape(-1)[1:1 + 23 * 5].reshape(23, 5)            # misaligned base pointer (4-byte path)
    assert torch.equal(thg.gather_rows(off, idx[:2]), off[idx[:2]])
    assert thg.gather_rows(src, idx[:0]).shape == (0, 2, 3)
    with pytest.raises(thg.ReferencePanic):
        thg.gather_rows(src, torch.tensor([24], device="cuda"))
    with pytest.raises(thg.ReferencePanic):
        thg.gather_rows(src, torch.tensor([-1], device="cuda"))
    with pytest.raises(ValueError):
        thg.gather_rows(src.cpu(), idx)
    with pytest.raises(ValueError):
        thg.gather_rows(src.transpose(0, 1), idx)


def test_sampler_output_feeds_the_gather(thg, fakedataset):
    """x[samples] and perm[edge_index] for a sampled tree: what filter_data does after the sampler."""
    ei, n = fakedataset
    ptrs, idx, perm = thg.to_csc(torch.from_numpy(ei).cuda(), n)
    x = torch.randn(n, 64, device="cuda")
    samples, rows, cols, eidx, _ = thg.neighbor_sampling_homogenous(ptrs, idx, torch.arange(32, device="cuda"), [4, 3])
    assert torch.equal(thg.gather_rows(x, samples), x[samples])
    orig_edges = thg.gather_rows(perm, eidx)       # positions in the caller's edge_index (quirk Q5)
    e = torch.from_numpy(ei).cuda()
    assert torch.equal(e[0][orig_edges], samples[rows]) and torch.equal(e[1][orig_edges], samples[cols])


def test_packed_host_transfer_of_sampled_batches(thg):
    """SampledBatches.to_host: the packed D2H path of the end-to-end benchmark returns exactly the per-batch results
    (rows comes from the cached host arange)."""
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "fakedataset.npz"))
    ei, n = torch.as_tensor(d["edge_index"]).cuda(), int(d["num_nodes"])
    ptrs, idx, _ = thg.to_csc(ei, n)
    B, S = 7, 33
    seeds = torch.randint(0, n, (B, S), device="cuda")
    plan = thg.HomogenousSampler(ptrs, idx, B, S, [6, 4, 2])
    res = plan.sample(seeds, seed=3)
    call = plan._call
    host = thg.HostBatches(4, int(call.cap_n[0]), int(call.cap_e[0]), S, seeds.device, fill=1.0)
    for first, count in ((0, 4), (4, 3)):
        nbytes = res.to_host(host, first, count)
        torch.cuda.synchronize()
        assert nbytes == 8 * int(res.samples_len[first:first + count].sum() + 2 * res.edges_len[first:first + count].sum())
        for i in range(count):
            want = res.batch(first + i)
            got = host.batch(i)
            for g, w in zip(got, want[:4]):
                assert torch.equal(g, w.cpu())


@pytest.mark.parametrize("transport,threads", [("compact", 1), ("compact", 3), ("hybrid", 2)])
def test_compact_host_transport_lands_the_same_vectors(thg, transport, threads):
    """transport="compact": i32 ids + u8 per-node edge counts on the bus, i64 vectors rebuilt by host threads
    (tchgeo_pack_transport / tchgeo_host_unpack_transport) -- identical to the plain packed copy, batch by batch,
    including batches whose nodes draw no edge and a temporal filter that thins the runs."""
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "fakedataset.npz"))
    ei, n = torch.as_tensor(d["edge_index"]).cuda(), int(d["num_nodes"])
    ptrs, idx, _ = thg.to_csc(ei, n)
    B, S = 9, 40
    seeds = torch.randint(0, n, (B, S), device="cuda")
    ts = torch.randint(0, 100, (idx.numel(),), device="cuda")
    flt = thg.TemporalEdgeFilter((0, 30), ts, True, thg.TEMPORAL_SAMPLE_STATIC)
    w = torch.rand(idx.numel(), dtype=torch.float64, device="cuda") + 0.1
    for plan, kw in ((thg.HomogenousSampler(ptrs, idx, B, S, [7, 5, 3]), {}),
                     (thg.HomogenousSampler(ptrs, idx, B, S, [4, 9], sampler=thg.UniformEdgeSampler(True)), {}),
                     (thg.HomogenousSampler(ptrs, idx, B, S, [6, 4], sampler=thg.WeightedEdgeSampler(w)), {}),
                     (thg.HomogenousSampler(ptrs, idx, B, S, [7, 5], filter=flt),
                      {"inputs_state": torch.zeros((B, S), dtype=torch.int64, device="cuda")})):
        res = plan.sample(seeds, seed=5, **kw)
        call = plan._call
        host = thg.HostBatches(5, int(call.cap_n[0]), int(call.cap_e[0]), S, seeds.device, fill=1.0, transport=transport,
                               threads=threads)
        for first, count in ((0, 5), (5, 4), (2, 1)):
            nbytes = res.to_host(host, first, count)
            ns, ne = int(res.samples_len[first:first + count].sum()), int(res.edges_len[first:first + count].sum())
            assert nbytes == 5 * ns + (8 if transport == "hybrid" else 4) * ne + 4
            for i in range(count):
                want = res.batch(first + i)
                got = host.batch(i)
                for g, w in zip(got, want[:4]):
                    assert g.dtype == torch.int64 and torch.equal(g, w.cpu())


def test_compact_host_transport_refuses_what_it_cannot_carry(thg):
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "fakedataset.npz"))
    ei, n = torch.as_tensor(d["edge_index"]).cuda(), int(d["num_nodes"])
    ptrs, idx, _ = thg.to_csc(ei, n)
    seeds = torch.randint(0, n, (2, 8), device="cuda")
    plan = thg.HomogenousSampler(ptrs, idx, 2, 8, [300])
    res = plan.sample(seeds, seed=1)
    host = thg.HostBatches(2, int(plan._call.cap_n[0]), int(plan._call.cap_e[0]), 8, seeds.device, fill=1.0,
                           transport="compact")
    with pytest.raises(ValueError):
        res.to_host(host)
