"""One short process for a GPU box with little time: the partition form of to_csc / to_csr against the radix form
(bit-exact, every column-length class: <= 32, <= 64, <= 1024, above, and a column that forces the hand-over), then the
time of both forms on the products-shaped graph.  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tch-geometric_b200")]
import tch_geometric as thg  # noqa: E402
from tools import synth      # noqa: E402


def form(f, fn, ei, size):
    os.environ["TCHGEO_CSX_SORT"] = f
    return fn(ei, size)


def main():
    dev = "cuda:0"
    out = {"equal": {}, "ms": {}}
    ei, n = synth.products_like(dev)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    hub = torch.stack([torch.randint(0, 50000, (40000,), generator=g, device=dev), torch.zeros(40000, dtype=torch.int64, device=dev)])
    hub[1, :3000] = torch.randint(0, 1000, (3000,), generator=g, device=dev)
    heavy = torch.stack([torch.randint(0, 50000, (60000,), generator=g, device=dev), torch.randint(0, 6, (60000,), generator=g, device=dev) * 100])
    mid = torch.stack([torch.randint(0, 100000, (400000,), generator=g, device=dev), torch.randint(0, 3000, (400000,), generator=g, device=dev)])
    cases = {"products_csc": (thg.to_csc, ei, n), "products_csr": (thg.to_csr, ei, n), "hub": (thg.to_csc, hub, (50000, 1000)),
             "heavy": (thg.to_csc, heavy, (50000, 600)), "columns_of_133": (thg.to_csc, mid, (100000, 3000))}
    for name, (fn, e_, size) in cases.items():
        a, b = form("cub", fn, e_, size), form("partition", fn, e_, size)
        out["equal"][name] = all(torch.equal(x, y) for x, y in zip(a, b))
        del a, b
    for name in ("products_csc", "products_csr"):
        fn, e_, size = cases[name]
        out["ms"][name] = {}
        for f in ("cub", "partition"):
            form(f, fn, e_, size)
            ms = []
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                form(f, fn, e_, size)
                b.record()
                torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            out["ms"][name][f] = float(np.median(ms))
    out["all_equal"] = all(out["equal"].values())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
