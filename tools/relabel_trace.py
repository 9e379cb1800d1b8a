"""Decode a TCHGEO_RELABEL_TRACE dump (csrc/relabel.cu, debug): per-CTA globaltimer stamps at the phase boundaries of the
persistent relabel kernel -> mean / max duration of every phase and of every wait, over the CTAs of group 0.

    TCHGEO_RELABEL_TRACE=/tmp/t.bin python bench.py --workload relabel --steps 1 --warmup 1; python tools/relabel_trace.py /tmp/t.bin
"""
import sys

import numpy as np

raw = open(sys.argv[1], "rb").read()
grid, T, groups, C = np.frombuffer(raw[:16], dtype=np.int32)
t = np.frombuffer(raw[16:], dtype=np.uint64).reshape(grid, T).astype(np.int64)
t = t[:C]                                   # group 0
t0 = t[:, 0].min()
names = ["start", "clear x2 done"]
# per pair: for tb in 0,1: (wait_clear_end, insert_end); for tb: (wait_insert_end, compact_end); for tb: (wait_compact_end, clear_end)
per_pair = []
for ph in ("insert", "compact", "clear"):
    for tb in (0, 1):
        per_pair += [f"wait before {ph}[{tb}]", f"{ph}[{tb}] done"]
k = 2
pair = 0
while k + len(per_pair) <= T and (t[:, k + len(per_pair) - 1] > 0).all():
    names += [f"pair {pair}: {x}" for x in per_pair]
    k += len(per_pair)
    pair += 1
print(f"grid {grid}, {groups} groups x {C} CTAs; {pair} full pairs traced; all times in us relative to the first CTA's start")
prev = t[:, 0]
for i, nm in enumerate(names[1:], start=1):
    cur = t[:, i]
    d = (cur - prev) / 1e3
    print(f"{nm:32s} reached at {((cur.min() - t0) / 1e3):9.1f} .. {((cur.max() - t0) / 1e3):9.1f}   step mean {d.mean():7.1f} max {d.max():7.1f} min {d.min():7.1f}")
    prev = cur
