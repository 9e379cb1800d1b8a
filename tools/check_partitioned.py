"""Multi-GPU parity check of the range-partitioned CSC path at the FULL BASELINE config-5 shape.

    torchrun --nproc-per-node G tools/check_partitioned.py [--scale 1.0] [--batches 32] [--protocol fixed|legacy]

Every rank builds its share of the papers100M-shaped graph (13.9 M columns / 202 M edges per rank at scale 1: with 8
ranks exactly N = 111 059 956, E = 1 615 685 872), the shares are gathered so that every rank ALSO holds the whole CSC
(13.8 GB: it fits one B200), and the partitioned sampler (frontier exchange over NVLink, this rank's own seed batches)
is compared with the replicated sampler on the same seeds bit for bit: samples, rows, cols, edge_index, layer offsets of
every batch.  Prints one JSON line from rank 0; exit code 1 on a mismatch."""
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
from tch_geometric.partitioned import DistComm, PartitionedPlan, PartitionedPlanF, PartitionedPlanGroups  # noqa: E402
from tools import synth  # noqa: E402


def gather_full_csc(part, world, rank, device):
    """every rank's (rebased colptr, row_indices) -> the whole CSC on every rank (one broadcast per owner)"""
    ptr_parts, idx_parts = [], []
    for r in range(world):
        meta = torch.tensor([part.ptrs.numel(), part.indices.numel(), part.edge_base], dtype=torch.int64, device=device)
        dist.broadcast(meta, src=r)
        np_, ni_, base = (int(x) for x in meta.tolist())
        p = part.ptrs.clone() if r == rank else torch.empty(np_, dtype=torch.int64, device=device)
        i = part.indices if r == rank else torch.empty(ni_, dtype=torch.int64, device=device)
        dist.broadcast(p, src=r)
        dist.broadcast(i, src=r)
        ptr_parts.append((p[:-1] if r < world - 1 else p) + base)
        idx_parts.append(i)
    return torch.cat(ptr_parts), torch.cat(idx_parts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--batches", type=int, default=32)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--protocol", default="fixed", choices=["fixed", "legacy"])
    ap.add_argument("--slack", type=float, default=1.5)
    ap.add_argument("--groups", type=int, default=1, help="fixed protocol: batch groups pipelined on separate streams")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=device)
    t0 = time.time()
    part, n, e_total, cols_rank = synth.papers_partition(thg, rank, world, device, args.scale)
    ptrs, idx = gather_full_csc(part, world, rank, device)
    assert ptrs.numel() == n + 1 and idx.numel() == e_total and int(ptrs[-1].item()) == e_total
    torch.cuda.synchronize()
    t_build = time.time() - t0
    fan, S, B = [15, 10, 5], 1024, args.batches
    if args.protocol == "fixed" and args.groups > 1:
        plan = PartitionedPlanGroups(part, B, S, fan, comm=DistComm(), groups=args.groups, slack=args.slack)
    elif args.protocol == "fixed":
        plan = PartitionedPlanF(part, B, S, fan, comm=DistComm(), slack=args.slack)
    else:
        plan = PartitionedPlan(part, B, S, fan, comm=DistComm())
    ref = thg.HomogenousSampler(ptrs, idx, B, S, fan)
    ok, edges, mism = True, 0, []
    for step in range(args.steps):          # twice: the second step reuses the exchange buffers
        bb = (step * world + rank) * B
        seeds = torch.from_numpy(synth.seed_batches(n, B, S, first_batch=bb)).to(device)
        got = plan.sample(seeds, seed=31 + step, batch_base=bb)
        want = ref.sample(seeds, seed=31 + step, batch_base=bb)
        same = bool((got.samples_len == want.samples_len).all() and (got.edges_len == want.edges_len).all()
                    and (got.layer_offsets == want.layer_offsets).all())
        for b in range(B):
            if not same:
                break
            same &= all(torch.equal(g, x) for g, x in zip(got.batch(b)[:4], want.batch(b)[:4]))
        if not same:
            mism.append(step)
        ok &= same
        edges += int(got.edges_len.sum())
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    tot = torch.tensor([edges], dtype=torch.int64, device=device)
    dist.all_reduce(tot)
    if rank == 0:
        print(json.dumps({"check": "partitioned == replicated, bit for bit, every batch of every rank",
                          "ok": bool(flag.item()), "protocol": args.protocol, "world": world, "N": n, "E": e_total,
                          "columns_per_rank": cols_rank, "batches_per_rank": B, "steps": args.steps,
                          "edges_compared": int(tot.item()), "rank0_mismatching_steps": mism,
                          "full_csc_bytes_per_rank": int(ptrs.numel() + idx.numel()) * 8, "build_s": round(t_build, 1),
                          "slack": args.slack if args.protocol == "fixed" else None,
                          "pipelined_batch_groups": args.groups if args.protocol == "fixed" else None}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
