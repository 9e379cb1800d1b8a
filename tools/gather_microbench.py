"""Random 8-byte gather ceiling on this GPU (context for the sampling roofline): torch.index_select over a
495 MB int64 table with uniform random indices, with and without the 32-byte L2 fetch granularity hint."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tch-geometric_b200"))
import torch  # noqa: E402

import tch_geometric as thg  # noqa: E402


def run(tag):
    n, m = 61_859_140, 123_900_000
    a = torch.arange(n, dtype=torch.int64, device="cuda")
    r = torch.randint(0, n, (m,), device="cuda")
    out = torch.empty(m, dtype=torch.int64, device="cuda")
    for _ in range(3):
        torch.index_select(a, 0, r, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.index_select(a, 0, r, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{tag}: {ms:.3f} ms per {m / 1e6:.1f}M gathers -> {m / ms / 1e6:.2f} G gathers/s; "
          f"sector-level {m * (8 + 32 + 8) / ms / 1e6:.0f} GB/s, algorithmic {m * 24 / ms / 1e6:.0f} GB/s")
    # sorted indices for contrast (fully coalesced)
    rs = torch.sort(r).values
    e0.record()
    for _ in range(10):
        torch.index_select(a, 0, rs, out=out)
    e1.record()
    torch.cuda.synchronize()
    print(f"{tag} sorted idx: {e0.elapsed_time(e1) / 10:.3f} ms")


if __name__ == "__main__":
    run("default")
    print("granularity now", thg.set_l2_fetch_granularity(32))
    run("l2fetch32")
