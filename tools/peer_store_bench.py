"""What do stores into NVLink peer memory cost per access pattern?  (background for the partitioned exchange)

    torchrun --nproc-per-node G tools/peer_store_bench.py [--mb 256]

Every rank owns one buffer in torch symmetric memory and writes into the buffer of rank (r + 1) % G -- or, with
--all, an equal share into every peer's buffer -- with plain torch ops on tensor views of the peer memory
(`_SymmetricMemory.get_buffer`), so no kernel of this repository is involved:
  contiguous   peer.copy_(local)                       whole lines, what part_put_kernel / the staged answer rows do
  rows16       peer[perm] = local   (16-byte rows)     one 16-byte store per request row, arbitrary order
  rows40_half  peer[:, :5] = a; peer[:, 5:] = b        an answer row written as two interleaved 20-byte halves
Prints one JSON line with GB/s per pattern (CUDA events, max over ranks).  NOT part of the product path; the first
8-GPU measurement of the peer exchange (row-by-row stores: 9.07 ms per step against 8.5 ms with NCCL) is what asks
for these numbers.  Written at the end of round 1, after the GPU budget was spent: not run yet."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def timed(fn, iters, device):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=device, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256, help="bytes written per rank per iteration (MiB)")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--all", action="store_true", help="spread the writes over all peers instead of one neighbour")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=device)
    import torch.distributed._symmetric_memory as symm
    nbytes = args.mb << 20
    buf = symm.empty(nbytes // 8, dtype=torch.int64, device=device)
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    peers = [p for p in range(world) if p != rank] if args.all else [(rank + 1) % world]
    share = nbytes // len(peers)
    res = {}

    def slot(p):  # which share of peer p's buffer this rank writes (writers of p are all ranks but p)
        return (rank if rank < p else rank - 1) if args.all else 0

    # contiguous
    src = torch.arange(share // 8, dtype=torch.int64, device=device)
    views = [hdl.get_buffer(p, (share // 8,), torch.int64, slot(p) * (share // 8)) for p in peers]

    def contiguous():
        for v in views:
            v.copy_(src)
    res["contiguous"] = contiguous

    # 16-byte rows in arbitrary order
    rows = share // 16
    src16 = torch.arange(rows * 2, dtype=torch.int64, device=device).view(rows, 2)
    perm = torch.randperm(rows, device=device)
    views16 = [hdl.get_buffer(p, (rows, 2), torch.int64, slot(p) * rows * 2) for p in peers]

    def rows16():
        for v in views16:
            v.index_copy_(0, perm, src16)
    res["rows16"] = rows16

    # 40-byte rows written as two 20-byte halves
    rows40 = share // 40
    a = torch.ones((rows40, 5), dtype=torch.int32, device=device)
    views40 = [hdl.get_buffer(p, (rows40, 10), torch.int32, slot(p) * rows40 * 10) for p in peers]

    def rows40_half():
        for v in views40:
            v[:, :5] = a
            v[:, 5:] = a
    res["rows40_half"] = rows40_half

    out = {"world": world, "bytes_per_rank_per_iter": nbytes, "peers_per_rank": len(peers), "GBps": {}}
    for name, fn in res.items():
        ms = timed(fn, args.iters, device)
        written = share * len(peers) if name != "rows40_half" else rows40 * 40 * len(peers)
        out["GBps"][name] = written / (ms * 1e-3) / 1e9
        hdl.barrier(channel=0)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
