"""Test infrastructure: the partitioned-CSC protocol written with plain torch ops (device-agnostic).  The 2-rank gloo
tests run it on the CPU with the oracle as the owner side, the GPU tests use it as a cross-check of the CUDA pipeline.
Not part of the product package."""
from typing import Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

from tch_geometric import _native as N
from tch_geometric.ops import _extract_sampler, _rng_get
from tch_geometric.partitioned import ColumnPartition, DistComm, SingleComm, cuda_serve


class PartitionedSampler:
    """neighbor_sampling_homogenous over a column-partitioned CSC.  `sample` is collective: every rank of
    the communicator must call it (with its own seed batches) the same number of times."""

    def __init__(self, part: ColumnPartition, num_neighbors: Sequence[int], sampler=None, comm=None, serve=None):
        self.part = part
        self.fanouts = [int(k) for k in num_neighbors]
        self.kind, w = _extract_sampler(sampler, hetero=False)
        if self.kind == N.SAMPLER_WEIGHTED and part.weights is None:
            raise ValueError("weighted sampling needs ColumnPartition.weights_local")
        self.comm = comm if comm is not None else (DistComm() if dist.is_available() and dist.is_initialized() else SingleComm())
        self.serve = serve if serve is not None else cuda_serve
        self.stats = {"requests_sent": 0, "request_bytes": 0, "answer_bytes": 0}

    def sample(self, inputs: Tensor, seed: Optional[int] = None, batch_base: int = 0):
        """inputs [B, S] (this rank's batches) -> list of B tuples (samples, rows, cols, edge_index, layer_offsets)."""
        if inputs.dim() != 2 or inputs.dtype != torch.int64:
            raise ValueError("inputs must be an int64 tensor of shape [B, S]")
        seed = _rng_get() if seed is None else seed
        dev, (B, S), world = inputs.device, inputs.shape, self.comm.world
        C = self.part.cols_per_rank
        ar = lambda n: torch.arange(n, dtype=torch.int64, device=dev)
        ids = inputs.reshape(-1)
        bidx = ar(B).repeat_interleave(S)
        pos = ar(S).repeat(B)
        front = torch.full((B,), S, dtype=torch.int64, device=dev)      # frontier size per batch (batch-major order)
        node_len = front.clone()
        edge_len = torch.zeros(B, dtype=torch.int64, device=dev)
        key_dtype = torch.uint8 if world <= 255 else torch.int16         # few-bit sort keys: one radix pass
        hop_ids, hop_cols, hop_eidx, hop_cnt, hop_lo = [], [], [], [], []
        for k in self.fanouts:
            hop_lo.append(torch.stack([node_len, edge_len, node_len], dim=1))
            F = ids.numel()
            if world > 1:
                owner = torch.clamp(torch.div(ids, C, rounding_mode="floor"), 0, world - 1)
                order = torch.sort(owner.to(key_dtype), stable=True).indices
                send_counts = torch.bincount(owner, minlength=world)
                req_ids = ids[order]
                req_meta = ((bidx[order] + batch_base) << 32) | pos[order]
            else:
                order, send_counts = None, torch.tensor([F], dtype=torch.int64, device=dev)
                req_ids, req_meta = ids, ((bidx + batch_base) << 32) | pos
            recv_counts, r_ids, r_meta = self.comm.exchange(send_counts, req_ids, req_meta)
            o_ids, o_ptrs = self.serve(self.part, r_ids, r_meta, k, self.kind, seed, 0)
            _, a_ids, a_ptrs = self.comm.exchange(recv_counts, o_ids, o_ptrs)
            self.stats["requests_sent"] += F
            self.stats["request_bytes"] += 16 * F
            self.stats["answer_bytes"] += 16 * k * F
            if order is not None:                                       # answers back in frontier order
                inv = torch.empty_like(order)
                inv[order] = ar(F)
                a_ids, a_ptrs = a_ids[inv], a_ptrs[inv]
            nz = (a_ptrs >= 0).nonzero()                                # row-major: frontier order, then slot order
            rows_nz = nz[:, 0]
            flat = rows_nz * max(k, 1) + nz[:, 1]
            new_ids, new_eidx = a_ids.reshape(-1)[flat], a_ptrs.reshape(-1)[flat]
            new_cols, new_bidx = pos[rows_nz], bidx[rows_nz]
            # per-batch counts by segment sums over the batch-major frontier (no atomics)
            csum = torch.cumsum((a_ptrs >= 0).sum(dim=1), 0)
            ends = torch.cumsum(front, 0)
            at_end = torch.where(ends > 0, csum[torch.clamp(ends - 1, min=0)], torch.zeros_like(ends)) if F > 0 \
                else torch.zeros_like(ends)
            per_batch = at_end - torch.cat([at_end.new_zeros(1), at_end[:-1]])
            first = torch.cumsum(per_batch, 0) - per_batch
            new_pos = ar(new_ids.numel()) - first[new_bidx] + node_len[new_bidx]
            hop_ids.append(new_ids); hop_cols.append(new_cols); hop_eidx.append(new_eidx); hop_cnt.append(per_batch)
            node_len = node_len + per_batch
            edge_len = edge_len + per_batch
            front = per_batch
            ids, bidx, pos = new_ids, new_bidx, new_pos
        # per-batch results: every hop's arrays are batch-major, so batch b's share of hop h is one slice
        H = len(self.fanouts)
        cnt_host = torch.stack(hop_cnt, 0).tolist() if H else []        # one host read of [H, B] counts
        lo_host = [x.tolist() for x in hop_lo]
        offs = [0] * H
        rows_all = torch.arange(S, S + (max((sum(cnt_host[h][b] for h in range(H)) for b in range(B)), default=0)),
                                dtype=torch.int64, device=dev)
        out = []
        for b in range(B):
            parts_s, parts_c, parts_e, e_b = [inputs[b]], [], [], 0
            for h in range(H):
                c = cnt_host[h][b]
                sl = slice(offs[h], offs[h] + c)
                parts_s.append(hop_ids[h][sl]); parts_c.append(hop_cols[h][sl]); parts_e.append(hop_eidx[h][sl])
                offs[h] += c
                e_b += c
            lo = [tuple(int(v) for v in lo_host[h][b]) for h in range(H)]
            empty = torch.zeros(0, dtype=torch.int64, device=dev)
            out.append((torch.cat(parts_s), rows_all[:e_b], torch.cat(parts_c) if parts_c else empty,
                        torch.cat(parts_e) if parts_e else empty, lo))
        return out
