"""Synthetic graphs shaped like the datasets BASELINE.json names (there is no network for the real
ones).  Generators are deterministic given the seed and run on any torch device, so the same code
builds the 62M-edge products-shaped graph on a B200 and a tiny one on the CPU for tests.

products-shaped: N = 2 449 029 nodes, E = 61 859 140 CSC entries, in-degrees
lognormal(sigma = 1.0) clipped to [1, 17 481] and rescaled to sum to E, sources uniform over [0, N)
without duplicates inside a column, generator seed 42.
"""
import numpy as np
import torch

PRODUCTS = dict(num_nodes=2_449_029, num_edges=61_859_140, max_degree=17_481)
# ogbn-mag-shaped (config 4): node counts and (src, rel, dst, edges)
MAG_NODES = dict(paper=736_389, author=1_134_649, institution=8_740, field_of_study=59_965)
MAG_RELS = [("author", "affiliated_with", "institution", 1_043_998), ("author", "writes", "paper", 7_145_660),
            ("paper", "cites", "paper", 5_416_271), ("paper", "has_topic", "field_of_study", 7_505_078)]


def lognormal_degrees(num_nodes, num_edges, sigma=1.0, dmin=1, dmax=17_481, seed=42):
    """Integer in-degrees with the requested sum: lognormal, clipped, rescaled, largest-remainder rounding."""
    rng = np.random.default_rng(seed)
    raw = rng.lognormal(0.0, sigma, num_nodes)
    dmax = min(dmax, num_edges)
    d = raw * (num_edges / raw.sum())
    for _ in range(50):
        d = np.clip(d, dmin, dmax)
        free = (d > dmin) & (d < dmax)
        excess = num_edges - d.sum()
        if abs(excess) < 0.5 or not free.any():
            break
        d[free] *= 1.0 + excess / d[free].sum()
    d = np.clip(d, dmin, dmax)
    fl = np.floor(d).astype(np.int64)
    rem = int(num_edges - fl.sum())
    if rem > 0:
        order = np.argsort(-(d - fl), kind="stable")
        cand = order[fl[order] < dmax][:rem]
        fl[cand] += 1
    elif rem < 0:
        order = np.argsort(d - fl, kind="stable")
        cand = order[fl[order] > dmin][:-rem]
        fl[cand] -= 1
    assert fl.sum() == num_edges, (fl.sum(), num_edges)
    return fl


def edges_from_degrees(deg, num_src, device, seed=42, shuffle=True):
    """COO edge_index [2, E] (row = src, col = dst) with deg[v] distinct uniform sources per column v."""
    deg_t = torch.as_tensor(deg, dtype=torch.int64, device=device)
    n_dst = deg_t.numel()
    E = int(deg_t.sum().item())
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    col = torch.repeat_interleave(torch.arange(n_dst, dtype=torch.int64, device=device), deg_t)
    row = torch.randint(0, num_src, (E,), generator=g, dtype=torch.int64, device=device)
    for _ in range(64):  # re-draw duplicates inside a column until none remain
        key = col * num_src + row
        skey, order = torch.sort(key)
        dup_sorted = torch.zeros(E, dtype=torch.bool, device=device)
        dup_sorted[1:] = skey[1:] == skey[:-1]
        nd = int(dup_sorted.sum().item())
        if nd == 0:
            break
        idx = order[dup_sorted]
        row[idx] = torch.randint(0, num_src, (nd,), generator=g, dtype=torch.int64, device=device)
        del key, skey, order, dup_sorted
    else:
        raise RuntimeError("could not remove duplicate edges")
    if shuffle:
        perm = torch.randperm(E, generator=g, device=device)
        row, col = row[perm], col[perm]
    return torch.stack([row, col])


def products_like(device, scale=1.0, seed=42):
    """-> (edge_index [2,E] on device, num_nodes).  scale < 1 shrinks N and E proportionally (tests)."""
    n = max(int(PRODUCTS["num_nodes"] * scale), 16)
    e = max(int(PRODUCTS["num_edges"] * scale), n)
    deg = lognormal_degrees(n, e, dmax=min(PRODUCTS["max_degree"], max(n // 4, 2)), seed=seed)
    return edges_from_degrees(deg, n, device, seed=seed), n


def mag_like(device, scale=1.0, seed=42):
    """-> (node_counts, {(src, rel, dst): edge_index})"""
    counts = {k: max(int(v * scale), 8) for k, v in MAG_NODES.items()}
    out = {}
    for i, (s, r, d, e) in enumerate(MAG_RELS):
        e = max(int(e * scale), counts[d])
        deg = lognormal_degrees(counts[d], e, dmin=0 if e < counts[d] else 1, dmax=max(counts[s] // 4, 2), seed=seed + i)
        out[(s, r, d)] = edges_from_degrees(deg, counts[s], device, seed=seed + i)
    return counts, out


def seed_batches(num_nodes, num_batches, seeds_per_batch, first_batch=0):
    """[num_batches, seeds_per_batch] distinct uniform node ids per batch, default_rng(1234 + b)."""
    out = np.empty((num_batches, seeds_per_batch), dtype=np.int64)
    for b in range(num_batches):
        out[b] = np.random.default_rng(1234 + first_batch + b).choice(num_nodes, seeds_per_batch, replace=False)
    return out


def papers_partition(thg, rank, world, device, scale=1.0):
    """This rank's share of the papers100M-shaped graph (BASELINE config 5): 13 882 495 columns / ~202 M edges per rank,
    so 8 ranks hold exactly the 111 059 956-node, 1 615 685 872-edge shape (weak scaling in the number of ranks).
    -> (ColumnPartition, n, e_total, cols_per_rank).  Collective when world > 1 (all-gather of the edge counts)."""
    import torch.distributed as dist
    from tch_geometric.partitioned import ColumnPartition, partition_bounds
    cols_full, edges_full = 111_059_956 // 8 + 1, 1_615_685_872 // 8
    cols_rank = max(int(cols_full * scale), 64)
    n = cols_rank * world
    b, e = partition_bounds(n, rank, world)
    deg = lognormal_degrees(e - b, max(int(edges_full * scale), e - b), dmax=17_481, seed=42 + rank)
    ei = edges_from_degrees(deg, n, device, seed=42 + rank)            # rows over all N nodes, cols local
    ptrs, idx, _ = thg.to_csc(ei, (n, e - b))
    del ei
    torch.cuda.empty_cache()
    e_local = torch.tensor([idx.numel()], dtype=torch.int64, device=device)
    if world > 1:
        allc = [torch.zeros_like(e_local) for _ in range(world)]
        dist.all_gather(allc, e_local)
        edge_base = int(sum(int(x.item()) for x in allc[:rank]))
        e_total = int(sum(int(x.item()) for x in allc))
    else:
        edge_base, e_total = 0, int(e_local.item())
    return ColumnPartition(ptrs, idx, n, rank, world, edge_base), n, e_total, cols_rank
