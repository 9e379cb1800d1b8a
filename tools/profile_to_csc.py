"""Timeline of one to_csc call on the products-shaped graph (torch.profiler / CUPTI, no kernel replay): every kernel and
memcpy with its start offset and duration, so that gaps between launches are visible.   python tools/profile_to_csc.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tch-geometric_b200"))
sys.path.insert(0, ROOT)
import tch_geometric as thg  # noqa: E402
from tools import synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    ei, n = synth.products_like(dev)
    for _ in range(3):
        out = thg.to_csc(ei, n)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    out = thg.to_csc(ei, n)
    ev1.record()
    torch.cuda.synchronize()
    print("events: %.3f ms" % ev0.elapsed_time(ev1))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        out = thg.to_csc(ei, n)
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
