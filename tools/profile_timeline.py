"""Timeline of one library call on the products-shaped graph (torch.profiler / CUPTI, no kernel replay): every kernel,
memset and memcpy with its start offset, duration and the gap before it, so that idle time between launches is visible.
    python tools/profile_timeline.py to_csc | tempo_walk"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tch-geometric_b200"))
sys.path.insert(0, ROOT)
import tch_geometric as thg  # noqa: E402
from tools import synth  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "to_csc"
    dev = torch.device("cuda", 0)
    ei, n = synth.products_like(dev)
    if what == "to_csc":
        call = lambda: thg.to_csc(ei, n)
    else:
        rp, ci, _ = thg.to_csr(ei, n)
        g = torch.Generator(device=dev)
        g.manual_seed(4242)
        ets = torch.randint(0, 1000, (ci.numel(),), generator=g, dtype=torch.int64, device=dev)
        nts = torch.randint(0, 1000, (n,), generator=g, dtype=torch.int64, device=dev)
        start = torch.arange(n, dtype=torch.int64, device=dev)
        sts = torch.randint(0, 1000, (n,), generator=g, dtype=torch.int64, device=dev)
        call = lambda: thg.tempo_random_walk(rp, ci, nts, ets, start, sts, 20, (0, 500), seed=7)
    for _ in range(3):
        out = call()
    torch.cuda.synchronize()
    import time
    for _ in range(3):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        out = call()
        ev1.record()
        torch.cuda.synchronize()
        print("events: %.3f ms, wall %.3f ms" % (ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        out = call()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    last_end = t0
    for e in evs:
        gap = e.time_range.start - last_end
        print("%9.1f us  +%8.1f us  gap %7.1f  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, gap, e.name[:90]))
        last_end = max(last_end, e.time_range.end)
    print("span %.1f us, kernel sum %.1f us" % (last_end - t0, sum(e.time_range.end - e.time_range.start for e in evs)))
    del out


if __name__ == "__main__":
    main()
