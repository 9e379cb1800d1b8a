"""all_to_all_single bandwidth probe (one rank per GPU): what the NCCL exchange of the partitioned path can reach.
    torchrun --nproc-per-node G tools/a2a_bench.py [--mb 1024]"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024, help="bytes sent per rank per call (MiB), split evenly over the ranks")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--pattern", default="static", choices=["static", "fresh", "uneven", "uneven_fresh_i32"],
                    help="static: same buffers, even splits; fresh: new output tensor per call; uneven: +-10% split sizes; "
                         "uneven_fresh_i32: what the partitioned path did (int32 rows of 40 B, counts exchange + host read)")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    n = a.mb * (1 << 20) // 8 // world * world
    src = torch.arange(n, dtype=torch.int64, device=dev)
    dst = torch.empty_like(src)
    for _ in range(3):
        dist.all_to_all_single(dst, src)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.Generator().manual_seed(rank)
    e0.record()
    for it in range(a.iters):
        if a.pattern == "static":
            dist.all_to_all_single(dst, src)
        elif a.pattern == "fresh":
            out = torch.empty(n + it * 4096, dtype=torch.int64, device=dev)[:n]
            dist.all_to_all_single(out, src)
        else:
            per = n // world
            sc = torch.tensor([per - int(x) for x in torch.randint(0, per // 10, (world,), generator=g)], device=dev)
            both = torch.empty((2, world), dtype=torch.int64, device=dev)
            both[0].copy_(sc)
            dist.all_to_all_single(both[1], both[0])
            scl, rcl = both.tolist()
            if a.pattern == "uneven":
                dist.all_to_all_single(dst[:sum(rcl)], src[:sum(scl)], output_split_sizes=rcl, input_split_sizes=scl)
            else:
                rows = src.view(torch.int32)[: sum(scl) * 10].view(-1, 10)[: sum(scl) // 5 * 5]
                scl5 = [x // 5 for x in scl]; rcl5 = [x // 5 for x in rcl]
                out = torch.empty((sum(rcl5), 10), dtype=torch.int32, device=dev)
                dist.all_to_all_single(out, rows[:sum(scl5)], output_split_sizes=rcl5, input_split_sizes=scl5)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    can = [torch.cuda.can_device_access_peer(local, j) for j in range(world) if j != local]
    if rank == 0:
        remote = n * 8 * (world - 1) / world
        print(json.dumps({"pattern": a.pattern, "world": world, "mib_per_rank": a.mb, "ms": t.item(), "remote_GBps_per_rank_per_dir": remote / t.item() / 1e6,
                          "peer_access_rank0": can, "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
