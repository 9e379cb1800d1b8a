"""Host-side rate of tchgeo_host_unpack_transport (no GPU): one group of 64 products-size batches, by thread count.
    python tools/host_unpack_bench.py            -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tch-geometric_b200"))
from tch_geometric import _native as N  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    B, nodes_per_batch = 64, 600_000
    counts = np.zeros((B, nodes_per_batch), np.uint8)
    counts[:, :115_000] = rng.integers(3, 7, (B, 115_000))          # ~5 edges per frontier node, leaves draw nothing
    nn = np.full(B, nodes_per_batch, np.int64)
    ne = counts.sum(axis=1).astype(np.int64)
    n_off = np.concatenate([[0], np.cumsum(nn)]).astype(np.int64)
    e_off = np.concatenate([[0], np.cumsum(ne)]).astype(np.int64)
    s32 = rng.integers(0, 2_400_000, n_off[-1]).astype(np.int32)
    e32 = rng.integers(0, 61_000_000, e_off[-1]).astype(np.int32)
    cnt = counts.reshape(-1)
    samples, cols, eidx = np.zeros(n_off[-1], np.int64), np.zeros(e_off[-1], np.int64), np.zeros(e_off[-1], np.int64)
    out_bytes = 8 * (n_off[-1] + 2 * e_off[-1])
    res = {}
    for simd, stores in (("", "nt"), ("", "regular"), ("sse2", "nt")):
        os.environ["TCHGEO_HOST_SIMD"] = simd
        os.environ["TCHGEO_HOST_STORES"] = stores
        for th in (1, 2, 4, 8, 12, 16, 32):
            if th > 2 * (os.cpu_count() or 1):
                continue
            best = 1e9
            for _ in range(3):
                t = time.perf_counter()
                st = N.lib.tchgeo_host_unpack_transport(s32.ctypes.data, e32.ctypes.data, cnt.ctypes.data, n_off.ctypes.data,
                                                        e_off.ctypes.data, B, samples.ctypes.data, cols.ctypes.data,
                                                        eidx.ctypes.data, th)
                best = min(best, time.perf_counter() - t)
                assert st == 0
            res[f"{simd or 'auto'}_{stores}_{th}"] = round(out_bytes / best / 1e9, 1)
    print(json.dumps({"what": "GB/s of i64 output written by tchgeo_host_unpack_transport, one group of 64 batches",
                      "out_GB": out_bytes / 1e9, "cpus": os.cpu_count(), "GBps_by_simd_threads": res}))


if __name__ == "__main__":
    main()
