"""Multi-GPU check + timing of the range-partitioned CSC path (BASELINE config 5 shape, scaled).

    torchrun --nproc-per-node G tools/check_partitioned.py [--scale 0.05] [--batches 16]

Every rank builds the same synthetic graph, keeps only its column range for sampling, and verifies that
the partitioned result (NCCL all-to-all frontier exchange) equals the replicated single-GPU sampler bit
for bit.  Prints one JSON line from rank 0."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tch-geometric_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import tch_geometric as thg  # noqa: E402
from tch_geometric.partitioned import ColumnPartition, DistComm, PartitionedPlan, PartitionedSampler  # noqa: E402
from tools import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--batches", type=int, default=16)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=device)
    ei, n = synth.products_like(device, scale=args.scale)
    ptrs, idx, _ = thg.to_csc(ei, n)
    part = ColumnPartition.from_full(ptrs, idx, rank, world)
    fan, S, B = [15, 10, 5], 1024, args.batches
    seeds = torch.from_numpy(synth.seed_batches(n, B, S, first_batch=rank * B)).to(device)
    ps = PartitionedSampler(part, fan, comm=DistComm())
    got = ps.sample(seeds, seed=31, batch_base=rank * B)
    want = thg.neighbor_sampling_homogenous_batched(ptrs, idx, seeds, fan, seed=31, batch_base=rank * B)
    ok = True
    for b in range(B):
        ok &= all(torch.equal(g, x) for g, x in zip(got[b][:4], want.batch(b)[:4]))
        ok &= list(got[b][4]) == list(want.batch(b)[4])
    edges = sum(int(g[1].numel()) for g in got)
    # the device pipeline (what bench.py --workload partitioned times): with the answer all-to-all, and with the answers
    # stored straight into the requesters' buffers over NVLink peer memory (the default when symmetric memory works)
    plan_a2a = PartitionedPlan(part, B, S, fan, comm=DistComm(), peer_answers=False)
    plan = PartitionedPlan(part, B, S, fan, comm=DistComm())
    peer_mode = plan.peer is not None
    for pl in (plan_a2a, plan):
        for sd in (31, 32):   # twice: the second call reuses the persistent buffers
            res = pl.sample(seeds, seed=sd, batch_base=rank * B)
            ref = want if sd == 31 else thg.neighbor_sampling_homogenous_batched(ptrs, idx, seeds, fan, seed=sd, batch_base=rank * B)
            ok_plan = bool((res.layer_offsets == ref.layer_offsets).all())
            for b in range(B):
                ok_plan &= all(torch.equal(g, x) for g, x in zip(res.batch(b)[:4], ref.batch(b)[:4]))
            ok &= ok_plan
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(args.iters):
        ps.sample(seeds, seed=100 + it, batch_base=rank * B)
    torch.cuda.synchronize()
    dist.barrier()
    dt_torch = (time.perf_counter() - t0) / args.iters
    t0 = time.perf_counter()
    for it in range(args.iters):
        plan_a2a.sample(seeds, seed=100 + it, batch_base=rank * B)
    torch.cuda.synchronize()
    dist.barrier()
    dt_a2a = (time.perf_counter() - t0) / args.iters
    t0 = time.perf_counter()
    for it in range(args.iters):
        plan.sample(seeds, seed=100 + it, batch_base=rank * B)
    torch.cuda.synchronize()
    dist.barrier()
    dt = (time.perf_counter() - t0) / args.iters
    flag = torch.tensor([1.0 if ok else 0.0, float(edges)], device=device, dtype=torch.float64)
    dist.all_reduce(flag, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"check": "partitioned == replicated (bit-exact)", "ranks_ok": int(flag[0].item()), "world": world,
                          "graph": {"nodes": n, "edges": int(idx.numel())}, "batches_per_rank": B,
                          "edges_per_call_all_ranks": flag[1].item(), "sec_per_call": dt,
                          "answer_exchange": "peer-memory stores from the serve kernel" if peer_mode else "all-to-all",
                          "sec_per_call_answer_all_to_all": dt_a2a,
                          "sec_per_call_torch_orchestration": dt_torch,
                          "edges_per_sec": flag[1].item() / dt, "request_bytes": ps.stats["request_bytes"],
                          "answer_bytes": ps.stats["answer_bytes"]}), flush=True)
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
