"""CPU-only, world_size 2 over gloo: the range-partitioned CSC path (config 5).  The all-to-all frontier
exchange and the requester-side layout run exactly as on the GPUs; the owner-side answers come from the
oracle's orc_serve_requests here (no GPU in this container).  The result must equal the single-process
reference-layout result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_serve(part, req_ids, req_meta, fanout, kind, seed, rel):
    from oracle import oracle as O
    sampler = {0: None, 1: ("uniform", True), 2: ("weighted", None if part.weights is None else part.weights.numpy())}[kind]
    ids, ptrs = O.serve_requests(part.ptrs.numpy(), part.indices.numpy(), part.col_begin, part.edge_base,
                                 req_ids.numpy(), req_meta.numpy(), fanout, sampler=sampler, seed=seed, rel=rel)
    return torch.from_numpy(ids), torch.from_numpy(ptrs)


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tch-geometric_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from tch_geometric.partitioned import ColumnPartition, DistComm
    from partitioned_reference import PartitionedSampler
    from tch_geometric import UniformEdgeSampler, WeightedEdgeSampler
    d = np.load(os.path.join(ROOT, "tests", "golden", "fakedataset.npz"))
    n = int(d["num_nodes"])
    ptrs, idx, _ = O.to_csc(d["edge_index"], n)
    w = np.random.default_rng(3).integers(1, 40, idx.size) / 8.0
    part = ColumnPartition.from_full(torch.from_numpy(ptrs), torch.from_numpy(idx), rank, world, torch.from_numpy(w))
    B, S, fan = 3, 17, [6, 4, 3]
    inputs = np.random.default_rng(100 + rank).integers(0, n, (B, S))
    ok = True
    for sampler, osamp in ((None, None), (UniformEdgeSampler(True), ("uniform", True)),
                           (WeightedEdgeSampler(torch.from_numpy(w)), ("weighted", w))):
        ps = PartitionedSampler(part, fan, sampler, comm=DistComm(), serve=_oracle_serve)
        res = ps.sample(torch.from_numpy(inputs), seed=55, batch_base=10 * rank)
        for b in range(B):
            want = O.neighbor_sampling_homogenous(ptrs, idx, inputs[b], fan, sampler=osamp, seed=55, batch=10 * rank + b)
            got = res[b]
            ok &= all((g.numpy() == x).all() and g.numel() == x.size for g, x in zip(got[:4], want[:4]))
            ok &= list(got[4]) == list(want[4])
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([ok, ps.stats["requests_sent"] > 0]))
    dist.barrier()
    dist.destroy_process_group()


def test_partitioned_two_ranks_match_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(os.path.join(str(tmp_path), f"ok{r}.npy")).all()


def test_partition_bounds_cover_all_columns():
    sys.path.insert(0, os.path.join(ROOT, "tch-geometric_b200"))
    from tch_geometric.partitioned import partition_bounds
    for n in (1, 7, 34, 1000):
        for world in (1, 2, 3, 8):
            b = [partition_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
