"""A/B of the two forms of tchgeo_coo_to_csx (TCHGEO_CSX_SORT=cub | partition) on one GPU: bit-exact comparison of
ptrs / indices / perm on the products-shaped graph (CSC and CSR), the mag-shaped relations and a few edge shapes, then
CUDA-event timings of both forms.  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tch-geometric_b200")]
import tch_geometric as thg  # noqa: E402
from tools import synth      # noqa: E402


def run(form, fn, ei, size):
    os.environ["TCHGEO_CSX_SORT"] = form
    return fn(ei, size)


def timed(form, fn, ei, size, reps=5):
    os.environ["TCHGEO_CSX_SORT"] = form
    for _ in range(2):
        fn(ei, size)
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(ei, size)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def main():
    dev = "cuda:0"
    out = {"equal": {}, "ms": {}}
    ei, n = synth.products_like(dev)
    cases = {"products_csc": (thg.to_csc, ei, n), "products_csr": (thg.to_csr, ei, n)}
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    for (s, r, d, e) in synth.MAG_RELS:
        ns, nd = synth.MAG_NODES[s], synth.MAG_NODES[d]
        rel = torch.stack([torch.randint(0, ns, (e,), generator=g, device=dev), torch.randint(0, nd, (e,), generator=g, device=dev)])
        cases["mag_" + r] = (thg.to_csc, rel, (ns, nd))
    star = torch.stack([torch.randint(0, 50000, (40000,), generator=g, device=dev), torch.zeros(40000, dtype=torch.int64, device=dev)])
    star[1, :3000] = torch.randint(0, 1000, (3000,), generator=g, device=dev)
    cases["hub_column_37000"] = (thg.to_csc, star, (50000, 1000))   # a column above PT_CAP: falls back
    mid = torch.stack([torch.randint(0, 50000, (60000,), generator=g, device=dev), torch.randint(0, 6, (60000,), generator=g, device=dev) * 100])
    cases["six_columns_of_10000"] = (thg.to_csc, mid, (50000, 600))   # columns above PT_WARP_MAX: sorted by the whole CTA
    for name, (fn, e_, size) in cases.items():
        a = run("cub", fn, e_, size)
        b = run("partition", fn, e_, size)
        out["equal"][name] = all(torch.equal(x, y) for x, y in zip(a, b))
        del a, b
    for name in ("products_csc", "products_csr", "mag_writes"):
        fn, e_, size = cases[name]
        out["ms"][name] = {f: timed(f, fn, e_, size) for f in ("cub", "partition")}
    out["all_equal"] = all(out["equal"].values())
    # per-kernel times of one partition-form call (CUPTI through torch.profiler, no replay)
    from torch.profiler import ProfilerActivity, profile
    os.environ["TCHGEO_CSX_SORT"] = "partition"
    fn, e_, size = cases["products_csc"]
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(e_, size)
        torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    out["partition_timeline_us"] = [[e.name.split("(")[0].split("::")[-1][:40], round(e.time_range.end - e.time_range.start, 1)] for e in evs]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
