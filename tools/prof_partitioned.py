import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tch-geometric_b200')
import torch
import tch_geometric as thg
from tch_geometric.partitioned import ColumnPartition, PartitionedSampler, SingleComm
from tools import synth
dev = torch.device('cuda', 0)
ei, n = synth.products_like(dev, scale=1.0)
ptrs, idx, _ = thg.to_csc(ei, n)
part = ColumnPartition.from_full(ptrs, idx, 0, 1)
B = 64
seeds = torch.from_numpy(synth.seed_batches(n, B, 1024)).to(dev)
ps = PartitionedSampler(part, [15, 10, 5], comm=SingleComm())
for _ in range(2): ps.sample(seeds, seed=1)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): out = ps.sample(seeds, seed=2)
torch.cuda.synchronize(); print('ms per call', (time.perf_counter() - t0) / 3 * 1e3, 'edges', sum(int(o[1].numel()) for o in out))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    ps.sample(seeds, seed=3); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
