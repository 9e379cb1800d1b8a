"""CPU-only, world_size 2 over gloo: the multi-GPU path shards seed batches / walkers over ranks with a
replicated graph and no data-path collective; only the job statistics are reduced (max time, sum units).
The oracle stands in for the device here (no GPU in this container); the GPU tests check that the device
honours the same global batch_base / walker_base indices."""
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


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tch-geometric_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from tch_geometric.sharding import reduce_job, shard_range
    d = np.load(os.path.join(ROOT, "tests", "golden", "karate.npz"))
    ptrs, idx, _ = O.to_csc(d["edge_index"], int(d["num_nodes"]))
    rp, ci, _ = O.to_csr(d["edge_index"], int(d["num_nodes"]))
    B, S, fan = 9, 4, [5, 5]
    inputs = np.random.default_rng(0).integers(0, 34, (B, S))
    b0, b1 = shard_range(B, rank, world)
    edges = 0
    pieces = []
    for b in range(b0, b1):
        s, r, c, e, lo = O.neighbor_sampling_homogenous(ptrs, idx, inputs[b], fan, seed=77, batch=b)
        edges += len(r)
        pieces.append(e)
    start = np.arange(34)
    w0, w1 = shard_range(start.size, rank, world)
    walks = O.random_walk(rp, ci, start[w0:w1], 8, 1.0, 0.5, seed=5, walker_base=w0)
    t, u = reduce_job(10.0 * (rank + 1), float(edges))
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), t=t, u=u, walks=walks,
             eidx=np.concatenate(pieces) if pieces else np.zeros(0, dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle as O
    d = np.load(os.path.join(ROOT, "tests", "golden", "karate.npz"))
    ptrs, idx, _ = O.to_csc(d["edge_index"], int(d["num_nodes"]))
    rp, ci, _ = O.to_csr(d["edge_index"], int(d["num_nodes"]))
    inputs = np.random.default_rng(0).integers(0, 34, (9, 4))
    want_e = [O.neighbor_sampling_homogenous(ptrs, idx, inputs[b], [5, 5], seed=77, batch=b)[3] for b in range(9)]
    want_w = O.random_walk(rp, ci, np.arange(34), 8, 1.0, 0.5, seed=5)
    r = [np.load(os.path.join(str(tmp_path), f"r{k}.npz")) for k in range(world)]
    assert (np.concatenate([x["eidx"] for x in r]) == np.concatenate(want_e)).all()
    assert (np.concatenate([x["walks"] for x in r]) == want_w).all()
    for x in r:  # every rank sees the whole-job figures: max time over ranks, total units
        assert float(x["t"]) == 20.0
        assert float(x["u"]) == float(sum(len(e) for e in want_e))
