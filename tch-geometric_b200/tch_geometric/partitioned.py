"""Range-partitioned CSC sampling (BASELINE config 5): the graph's columns are split over ranks and the
hop frontiers are exchanged with an all-to-all (NCCL over NVLink on the GPUs).

Every rank samples its OWN seed batches; per hop it
  1. buckets its frontier by owning rank (owner(w) = w // cols_per_rank),
  2. all-to-all's the requests (node id, global batch index, position in the batch's samples vector),
  3. answers the requests it received with `tchgeo_serve_requests` (the CUDA kernel; draws use the same
     Philox counters as the replicated path),
  4. all-to-all's the answers back and lays them out in frontier order.
Because the counters depend only on (seed, batch, position, degree), the result equals the single-GPU
`neighbor_sampling_homogenous` result bit for bit, whatever the partitioning.

The reference has no counterpart (it is single-process); the tree layout produced here is the one of
src/algo/neighbor_sampling.rs:162-230.

Drivers (all produce the replicated sampler's padded [B, capacity] layout, SampledBatches-style):
  * PartitionedPlanF   -- the product path (csrc/partitioned_fixed.cu): both exchanges of a hop are NVLink peer-memory
                          stores into fixed per-pair segments, every count stays on the device, no host synchronisation
                          inside a step.  PartitionedPlanGroups pipelines several batch groups of it on separate streams.
  * PartitionedPlan    -- round 1's protocol (csrc/partitioned.cu): the count matrix is read by the host once per hop
                          and the rows travel by NCCL all-to-all(v) or by peer stores at the offsets the matrix gives.
The torch-op restatement of the protocol used by the CPU (gloo) tests lives in tests/partitioned_reference.py.
"""
import ctypes
import os
import warnings
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist
from torch import Tensor

from . import _native as N
from .ops import _check, _extract_sampler, _ptr, _rng_get, _stream


def cols_per_rank(num_nodes: int, world: int) -> int:
    return max((num_nodes + world - 1) // world, 1)


def partition_bounds(num_nodes: int, rank: int, world: int) -> Tuple[int, int]:
    c = cols_per_rank(num_nodes, world)
    return min(rank * c, num_nodes), min((rank + 1) * c, num_nodes)


class ColumnPartition:
    """This rank's share of a CSC: columns [col_begin, col_end) with a colptr rebased to the local
    row_indices slice; edge_base = number of CSC entries owned by lower ranks."""

    def __init__(self, col_ptrs_local: Tensor, row_indices_local: Tensor, num_nodes: int, rank: int, world: int,
                 edge_base: int, weights_local: Optional[Tensor] = None):
        self.col_begin, self.col_end = partition_bounds(num_nodes, rank, world)
        if col_ptrs_local.numel() != self.col_end - self.col_begin + 1:
            raise ValueError("col_ptrs_local must have one entry per owned column plus one")
        self.ptrs, self.indices, self.weights = col_ptrs_local, row_indices_local, weights_local
        self.num_nodes, self.rank, self.world, self.edge_base = int(num_nodes), rank, world, int(edge_base)
        self.cols_per_rank = cols_per_rank(num_nodes, world)

    @staticmethod
    def from_full(col_ptrs: Tensor, row_indices: Tensor, rank: int, world: int, weights: Optional[Tensor] = None):
        """Slice a replicated CSC (tests and small graphs)."""
        n = col_ptrs.numel() - 1
        b, e = partition_bounds(n, rank, world)
        lo, hi = int(col_ptrs[b].item()), int(col_ptrs[e].item())
        return ColumnPartition((col_ptrs[b:e + 1] - lo).contiguous(), row_indices[lo:hi].contiguous(), n, rank, world, lo,
                               None if weights is None else weights[lo:hi].contiguous())


class SingleComm:
    rank, world = 0, 1

    def exchange(self, send_counts: Tensor, *tensors):
        return (send_counts,) + tuple(tensors)

    def exchange_rows(self, send_counts: Tensor, rows: Tensor, alloc=None):
        """-> (recv_counts list, send_counts list, received rows); reads the counts on the host (one sync)"""
        sc = send_counts.tolist()
        return sc, sc, rows

    def return_rows(self, rows: Tensor, n_rows: int, n_back: int, send_counts, recv_counts, alloc=None):
        return rows

    def all_gather_int(self, value: int, device):
        return [int(value)]


class DistComm:
    """all-to-all(v) over a torch.distributed group (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def exchange(self, send_counts: Tensor, *tensors):
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        sc, rc = send_counts.tolist(), recv_counts.tolist()
        outs = []
        for t in tensors:
            out = torch.empty((sum(rc),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_to_all_single(out, t.contiguous(), output_split_sizes=rc, input_split_sizes=sc, group=self.group)
            outs.append(out)
        return (recv_counts,) + tuple(outs)

    def exchange_rows(self, send_counts: Tensor, rows: Tensor, alloc=None):
        """Requests: all-to-all of the per-owner counts, one host read of both count vectors, all-to-all(v) of the rows.
        alloc(n) -> [max(n,1), ...] output buffer (default: a fresh tensor)."""
        both = torch.empty((2, self.world), dtype=send_counts.dtype, device=send_counts.device)
        both[0].copy_(send_counts)
        dist.all_to_all_single(both[1], both[0], group=self.group)
        sc, rc = both.tolist()
        out = alloc(sum(rc)) if alloc is not None else \
            torch.empty((max(sum(rc), 1),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(out[:sum(rc)], rows[:sum(sc)], output_split_sizes=rc, input_split_sizes=sc, group=self.group)
        return rc, sc, out

    def all_gather_int(self, value: int, device):
        mine = torch.tensor([int(value)], dtype=torch.int64, device=device)
        out = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine, group=self.group)
        return [int(x.item()) for x in out]

    def return_rows(self, rows: Tensor, n_rows: int, n_back: int, send_counts, recv_counts, alloc=None):
        """Answers travel the reverse way: what was received is sent back, split sizes swapped."""
        out = alloc(n_back) if alloc is not None else \
            torch.empty((max(n_back, 1),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(out[:n_back], rows[:n_rows], output_split_sizes=list(send_counts),
                               input_split_sizes=list(recv_counts), group=self.group)
        return out


def cuda_serve(part: ColumnPartition, req_ids: Tensor, req_meta: Tensor, fanout: int, kind: int, seed: int, rel: int = 0):
    """Answer requests with the CUDA kernel behind tchgeo_serve_requests -> (ids [n,k], ptrs [n,k])."""
    dev = part.ptrs.device
    _check(part.ptrs, torch.int64, "col_ptrs_local")
    _check(part.indices, torch.int64, "row_indices_local", dev)
    n = req_ids.numel()
    out_ids = torch.empty((n, fanout), dtype=torch.int64, device=dev)
    out_ptrs = torch.empty((n, fanout), dtype=torch.int64, device=dev)
    scratch = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = N.lib.tchgeo_serve_requests(_ptr(part.ptrs), _ptr(part.indices), _ptr(part.weights), part.col_begin,
                                         part.col_end - part.col_begin, part.edge_base, _ptr(req_ids.contiguous()),
                                         _ptr(req_meta.contiguous()), n, int(fanout), kind, seed, rel, _ptr(out_ids),
                                         _ptr(out_ptrs), _ptr(scratch), _stream(dev))
    N.check(st)
    return out_ids, out_ptrs


def serve_rows(part: ColumnPartition, req: Tensor, n: int, fanout: int, kind: int, seed: int, ans: Tensor, err: Tensor,
               rel: int = 0):
    """Owner side on interleaved rows (asynchronous): req [n,2] i64 -> ans [n, 2*fanout] i32 (ids | local csc
    positions); errors are OR-ed into `err`."""
    dev = part.ptrs.device
    with torch.cuda.device(dev):
        N.check(N.lib.tchgeo_serve_requests_rows(_ptr(part.ptrs), _ptr(part.indices), _ptr(part.weights), part.col_begin,
                                                 part.col_end - part.col_begin, part.indices.numel(), _ptr(req), int(n),
                                                 int(fanout), kind, seed, rel, _ptr(ans), _ptr(err), _stream(dev)))


class PartitionedBatches:
    """Result of PartitionedPlan.sample: B reference-layout results in padded [B, capacity] buffers (same access
    pattern as ops.SampledBatches)."""

    def __init__(self, plan, node_len, edge_len):
        self.samples, self.rows, self.cols, self.edge_index = plan.samples, plan.rows, plan.cols, plan.eidx
        self._H = len(plan.fanouts)
        self._node_len, self._edge_len = node_len, edge_len         # host [H+1, B]
        self.samples_len, self.edges_len = node_len[-1], edge_len[-1]

    def __len__(self):
        return self.samples.shape[0]

    @property
    def layer_offsets(self):
        """[B, H, 3]: LayerOffset(len(samples), len(edges), len(samples)) at the start of every hop"""
        nl, el = self._node_len[:-1].T, self._edge_len[:-1].T
        return np.stack([nl, el, nl], axis=2)

    def batch(self, b):
        ns, ne = int(self.samples_len[b]), int(self.edges_len[b])
        lo = [(int(self._node_len[h, b]), int(self._edge_len[h, b]), int(self._node_len[h, b])) for h in range(self._H)]
        return self.samples[b, :ns], self.rows[b, :ne], self.cols[b, :ne], self.edge_index[b, :ne], lo

    def to_host(self, host, first: int = 0, count: Optional[int] = None) -> int:
        """packed D2H of batches [first, first + count), like ops.SampledBatches.to_host"""
        from .ops import packed_to_host
        B = self.samples.shape[0]
        return packed_to_host(self.samples, self.cols, self.edge_index, self.samples_len, self.edges_len,
                              self.samples.shape[1], self.rows.shape[1], self.samples.device, host, first,
                              B - first if count is None else int(count))


class _PlanGroup:
    """A contiguous range of a plan's batches with its own stream, request buffer and workspace: the groups of a
    plan advance through the hop phases in lock step on different streams, so one group's all-to-all overlaps the
    other group's kernels."""

    def __init__(self, plan, b0, b1):
        dev = plan.device
        i64 = dict(dtype=torch.int64, device=dev)
        self.b0, self.b1, self.B = b0, b1, b1 - b0
        self.stream = torch.cuda.Stream(device=dev) if plan.num_groups > 1 else None
        fmax = max(plan.capF) if plan.capF else 0
        self.req = torch.empty((max(self.B * fmax, 1), 2), **i64)
        self.counts = torch.zeros((2, max(plan.comm.world, 1)), **i64)   # [0] per-owner request counts, [1] cursor
        ws = max((N.lib.tchgeo_part_hop_workspace_bytes(self.B, c) for c in plan.capF), default=0)
        if plan.capF and ws == 0:
            raise ValueError("frontier too large for one call: use fewer batches per call")
        self.ws = torch.empty(max(int(ws), 1), dtype=torch.uint8, device=dev)
        self.samples, self.rows = plan.samples[b0:b1], plan.rows[b0:b1]
        self.cols, self.eidx = plan.cols[b0:b1], plan.eidx[b0:b1]
        self.hop = {}   # transient tensors / counts of the hop in flight
        self._bufs = {}

    def buf(self, name, rows, cols, dtype, device):
        """[rows, cols] view of a persistent buffer that only ever grows (x1.25): the exchange sees stable addresses
        and the hot loop never reaches the allocator once the sizes have settled."""
        need = max(int(rows), 1) * int(cols)
        t = self._bufs.get(name)
        if t is None or t.dtype != dtype or t.numel() < need:
            t = torch.empty(int(need * 1.25) + 16, dtype=dtype, device=device)
            self._bufs[name] = t
        return t[:need].view(max(int(rows), 1), int(cols))


def peer_offsets(C, me):
    """Where rank `me`'s rows go when the two all-to-all(v)s of a hop are replaced by direct stores.
    C[q][o] = number of requests rank q sends to owner o (the all-gathered count matrix).  Returns
      recv_counts[q]  requests `me` receives from q,
      send_counts[o]  requests `me` sends to o,
      req_row0[o]     first row of me's group in owner o's request buffer (an all-to-all delivers the groups of the
                      lower-ranked requesters first),
      ans_row0[q]     first row of me's answers in requester q's answer buffer (q's requests to lower-ranked owners
                      come first)."""
    world = len(C)
    sc = list(C[me])
    rc = [C[q][me] for q in range(world)]
    req_row0 = [sum(C[q][o] for q in range(me)) for o in range(world)]
    ans_row0 = [sum(C[q][:me]) for q in range(world)]
    return rc, sc, req_row0, ans_row0


class _PeerExchange:
    """Both exchanges of a hop fused into the kernels on either side of them (NVLink peer memory instead of all-to-alls).

    Every rank owns a persistent request buffer and a persistent answer buffer in torch symmetric memory; `buffer_ptrs`
    gives every rank the device address of every other rank's buffers.  Per hop the ranks all-gather the [world, world]
    matrix of request counts (the hop's one host synchronisation), which tells every rank at which row of which peer
    buffer each of its rows belongs -- the row the all-to-all would have delivered it to:
      * tchgeo_part_scatter_hop stores every request row straight into the OWNER's request buffer,
      * tchgeo_serve_requests_rows_peer stores every answer row straight into the REQUESTER's answer buffer,
    and a symmetric-memory barrier on the stream follows each of the two kernels (requests landed -> serve; answers
    landed -> layout).  Those two barriers also order the next round of stores after the previous round's readers, so
    one buffer of each kind per rank is enough.  The request buffer is sized for `slack` times the mean load; a hop in
    which some owner would receive more falls back to the request all-to-all (every rank sees the same matrix, so all
    of them take the same branch)."""

    def __init__(self, plan, slack=2.0):
        import torch.distributed._symmetric_memory as symm
        comm, dev = plan.comm, plan.device
        group = comm.group if comm.group is not None else dist.group.WORLD
        words = max((plan.B * f * 2 * k for f, k in zip(plan.capF, plan.fanouts)), default=1)
        self.ans = symm.empty(max(int(words), 1), dtype=torch.int32, device=dev)
        self.ans_hdl = symm.rendezvous(self.ans, group)
        self.req_rows = int(slack * plan.B * max(plan.capF)) + 1024
        self.req = symm.empty(self.req_rows * 2, dtype=torch.int64, device=dev)
        self.req_hdl = symm.rendezvous(self.req, group)
        for hdl in (self.ans_hdl, self.req_hdl):
            if len(hdl.buffer_ptrs) != comm.world or int(hdl.rank) != comm.rank:
                raise RuntimeError("symmetric memory rendezvous does not match the communicator")
        self.ans_ptrs = np.array([int(x) for x in self.ans_hdl.buffer_ptrs], dtype=np.uint64)
        self.req_ptrs = np.array([int(x) for x in self.req_hdl.buffer_ptrs], dtype=np.uint64)
        self.cmat = torch.zeros((comm.world, comm.world), dtype=torch.int64, device=dev)
        self.group = group
        self.fallback_hops = 0

    def count_matrix(self, comm, counts):
        """-> (recv_counts, send_counts, row0 of my requests in every owner's buffer, row0 of my answers in every
        requester's buffer, whether every owner's load fits its request buffer)"""
        dist.all_gather_into_tensor(self.cmat.reshape(-1), counts.contiguous(), group=self.group)
        C = self.cmat.tolist()                       # the hop's one host synchronisation
        rc, sc, req_row0, ans_row0 = peer_offsets(C, comm.rank)
        fits = all(sum(C[q][o] for q in range(comm.world)) <= self.req_rows for o in range(comm.world))
        return rc, sc, req_row0, ans_row0, fits

    def requests_landed(self):
        self.req_hdl.barrier(channel=0)

    def answers_landed(self):
        self.ans_hdl.barrier(channel=0)


class PartitionedPlan:
    """neighbor_sampling_homogenous over a column-partitioned CSC, device pipeline.  `sample` is collective."""

    def __init__(self, part: ColumnPartition, num_batches: int, seeds_per_batch: int, num_neighbors: Sequence[int],
                 sampler=None, comm=None, serve_rows=None, edge_bases=None, groups: Optional[int] = None,
                 peer_answers: Optional[bool] = None):
        """serve_rows(r_req [n,2], recv_counts, fanout, seed, ans [n,2k] int32) overrides the owner side (tests simulate
        several owners on one GPU with it); default: tchgeo_serve_requests_rows over `part`.
        edge_bases: every rank's ColumnPartition.edge_base (default: all-gathered through the communicator).
        groups: batch groups pipelined on separate streams (default 1: the overlap did not pay on 2 B200s).
        peer_answers: store requests and answers straight into the peers' buffers over NVLink peer memory instead of
        two all-to-alls (_PeerExchange).  None = when it applies (2 CUDA ranks -- more with TCHGEO_PEER_ANSWERS=1, never
        with TCHGEO_PEER_ANSWERS=0 --, one group, default owner side) and the symmetric-memory rendezvous succeeds;
        True = required."""
        self.part = part
        self.serve_rows = serve_rows
        self.fanouts = [int(k) for k in num_neighbors]
        self.kind, _ = _extract_sampler(sampler, hetero=False)
        if self.kind == N.SAMPLER_WEIGHTED and part.weights is None:
            raise ValueError("weighted sampling needs ColumnPartition.weights_local")
        self.comm = comm if comm is not None else (DistComm() if dist.is_available() and dist.is_initialized() else SingleComm())
        _check(part.ptrs, torch.int64, "col_ptrs_local")
        dev = part.ptrs.device
        _check(part.indices, torch.int64, "row_indices_local", dev)
        if part.num_nodes >= 2 ** 31 or part.indices.numel() >= 2 ** 31:
            raise ValueError("the compact answer rows need node ids and a rank's CSC share below 2^31")
        self.device, self.B, self.S = dev, int(num_batches), int(seeds_per_batch)
        B, S, H = self.B, self.S, len(self.fanouts)
        # worst-case frontier / output sizes per batch (the recurrence of tchgeo_neighbor_sampling_capacity)
        self.capF, cap_e, f = [], 0, S
        for k in self.fanouts:
            self.capF.append(f)
            f *= k
            cap_e += f
        self.cap_n, self.cap_e = S + cap_e, max(cap_e, 1)
        i64 = dict(dtype=torch.int64, device=dev)
        self.samples = torch.empty((B, self.cap_n), **i64)
        self.rows = torch.empty((B, self.cap_e), **i64)
        self.cols = torch.empty((B, self.cap_e), **i64)
        self.eidx = torch.empty((B, self.cap_e), **i64)
        self.lens = torch.zeros((2, H + 1, B), **i64)               # [0] node_len, [1] edge_len after h hops
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        if edge_bases is None:
            edge_bases = self.comm.all_gather_int(part.edge_base, dev)
        if len(edge_bases) != self.comm.world:
            raise ValueError("edge_bases must have one entry per rank")
        self.edge_bases = torch.tensor([int(x) for x in edge_bases], **i64)
        if groups is None:
            groups = 1   # measured on 2 B200s: pipelining batch groups over streams is slower (7.7 ms -> 8.3 ms)
        self.num_groups = max(1, min(int(groups), max(B, 1)))
        cuts = [B * g // self.num_groups for g in range(self.num_groups + 1)]
        with torch.cuda.device(dev):
            self.groups = [_PlanGroup(self, cuts[g], cuts[g + 1]) for g in range(self.num_groups)]
        self.stats = {"requests_sent": 0, "request_bytes": 0, "answer_bytes": 0}
        self.profile = None   # set to {} to collect per-phase device times (ms, CUDA events, group 0's stream)
        self.peer = None
        applies = (isinstance(self.comm, DistComm) and self.comm.world > 1 and dev.type == "cuda"
                   and self.serve_rows is None and self.num_groups == 1 and H > 0)
        # measured default: on 2 B200s the peer-memory exchange wins (6.56 -> 4.58 ms per step); on 8 the fine-grained
        # remote stores (16-byte request rows, half-filled answer lines) lose to NCCL's bulk copies (9.1 vs 8.5 ms)
        env = os.environ.get("TCHGEO_PEER_ANSWERS")
        want = peer_answers if peer_answers is not None else (env != "0" and (env == "1" or self.comm.world <= 2))
        if peer_answers and not applies:
            raise ValueError("peer_answers needs several CUDA ranks, one batch group and the default owner side")
        if want and applies:
            # collective: every rank takes the same branch, and a failure is agreed on before anyone proceeds
            ok = 1
            try:
                with torch.cuda.device(dev):
                    self.peer = _PeerExchange(self)
            except Exception as e:  # noqa: BLE001  (no symmetric memory on this system: keep the all-to-all)
                ok, self.peer, err = 0, None, e
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.comm.group)
            if int(flag.item()) == 0:
                if peer_answers:
                    raise RuntimeError(f"symmetric-memory rendezvous failed on some rank{'' if ok else ': %r' % (err,)}")
                if not ok:
                    warnings.warn(f"peer-memory exchange unavailable, using the all-to-alls: {err!r}")
                self.peer = None

    def _mark(self, g, marks, name):
        if marks is not None and g is self.groups[0]:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    # ---- the five phases of a hop, each issued on the group's stream -------------------------------------
    def _begin(self, g, h, batch_base):
        lens = self.lens
        g.hop = {"fr_begin": lens[0, h - 1, g.b0:g.b1] if h > 0 else None}
        if self.peer is not None:
            # count -> all-gather of the count matrix -> scatter straight into the owners' request buffers
            common = (_ptr(g.samples), self.cap_n, _ptr(g.hop["fr_begin"]), _ptr(lens[0, h, g.b0:g.b1]), g.B, self.capF[h],
                      self.part.cols_per_rank, self.comm.world)
            N.check(N.lib.tchgeo_part_count_hop(*common, _ptr(g.counts[0]), _ptr(g.counts[1]), _ptr(self.err),
                                                _ptr(g.ws), g.ws.numel(), _stream(self.device)))
            rc, sc, req_row0, ans_row0, fits = self.peer.count_matrix(self.comm, g.counts[0])
            g.hop.update(rc=rc, sc=sc, row0=ans_row0, F=sum(sc), n_recv=sum(rc), peer_req=fits)
            row0 = np.array(req_row0, dtype=np.int64)
            N.check(N.lib.tchgeo_part_scatter_hop(*common, batch_base + g.b0, _ptr(g.counts[0]), _ptr(g.counts[1]),
                                                  _ptr(g.req), self.peer.req_ptrs.ctypes.data if fits else None,
                                                  row0.ctypes.data, _ptr(self.err), _ptr(g.ws), g.ws.numel(),
                                                  _stream(self.device)))
            return
        N.check(N.lib.tchgeo_part_begin_hop(_ptr(g.samples), self.cap_n, _ptr(g.hop["fr_begin"]), _ptr(lens[0, h, g.b0:g.b1]),
                                            g.B, self.capF[h], self.part.cols_per_rank, self.comm.world, batch_base + g.b0,
                                            _ptr(g.counts[0]), _ptr(g.counts[1]), _ptr(g.req), _ptr(self.err),
                                            _ptr(g.ws), g.ws.numel(), _stream(self.device)))

    def _requests(self, g):
        alloc = lambda n: g.buf("r_req", n, 2, torch.int64, self.device)
        if self.peer is not None:
            hp = g.hop
            if hp["peer_req"]:
                self.peer.requests_landed()          # every requester's stores are in every owner's buffer
                hp["r_req"] = self.peer.req[:max(hp["n_recv"], 1) * 2].view(max(hp["n_recv"], 1), 2)
            else:                                    # an owner's load exceeds its peer buffer: all-to-all for this hop
                self.peer.fallback_hops += 1
                out = alloc(hp["n_recv"])
                dist.all_to_all_single(out[:hp["n_recv"]], g.req[:hp["F"]], output_split_sizes=hp["rc"],
                                       input_split_sizes=hp["sc"], group=self.peer.group)
                hp["r_req"] = out
            return
        # host reads the counts: syncs g's stream only
        rc, sc, r_req = self.comm.exchange_rows(g.counts[0], g.req, alloc)
        g.hop.update(rc=rc, sc=sc, r_req=r_req, F=sum(sc), n_recv=sum(rc))

    def _serve(self, g, k, seed):
        hp = g.hop
        if self.peer is not None:
            part, dev = self.part, self.device
            rc = np.array(hp["rc"], dtype=np.int64)
            row0 = np.array(hp["row0"], dtype=np.int64)
            with torch.cuda.device(dev):
                N.check(N.lib.tchgeo_serve_requests_rows_peer(
                    _ptr(part.ptrs), _ptr(part.indices), _ptr(part.weights), part.col_begin, part.col_end - part.col_begin,
                    part.indices.numel(), _ptr(hp["r_req"]), int(hp["n_recv"]), int(k), self.kind, seed, 0, self.comm.world,
                    self.peer.ans_ptrs.ctypes.data, rc.ctypes.data, row0.ctypes.data, _ptr(self.err), _stream(dev)))
            return
        hp["ans"] = g.buf("ans", hp["n_recv"], 2 * k, torch.int32, self.device)
        if self.serve_rows is not None:
            self.serve_rows(hp["r_req"], hp["rc"], k, seed, hp["ans"])
        else:
            serve_rows(self.part, hp["r_req"], hp["n_recv"], k, self.kind, seed, hp["ans"], self.err)

    def _answers(self, g, k):
        hp = g.hop
        if self.peer is not None:
            self.peer.answers_landed()   # every owner's stores are in every requester's buffer
            hp["back"] = self.peer.ans[:max(hp["F"], 1) * 2 * k].view(max(hp["F"], 1), 2 * k)
            return
        hp["back"] = self.comm.return_rows(hp["ans"], hp["n_recv"], hp["F"], hp["sc"], hp["rc"],
                                           lambda n: g.buf("back", n, hp["ans"].shape[1], torch.int32, self.device))

    def _finish(self, g, h, k, batch_base):
        hp, lens, sl = g.hop, self.lens, slice(g.b0, g.b1)
        N.check(N.lib.tchgeo_part_finish_hop(_ptr(g.req), _ptr(hp["back"]), hp["F"], k, _ptr(self.edge_bases),
                                             self.part.cols_per_rank, self.comm.world,
                                             _ptr(hp["fr_begin"]), _ptr(lens[0, h, sl]), g.B, self.capF[h], _ptr(lens[0, h, sl]),
                                             _ptr(lens[1, h, sl]), _ptr(lens[0, h + 1, sl]), _ptr(lens[1, h + 1, sl]),
                                             _ptr(g.samples), self.cap_n, _ptr(g.rows), _ptr(g.cols), _ptr(g.eidx),
                                             self.cap_e, _ptr(self.err), _ptr(g.ws), g.ws.numel(), _stream(self.device)))
        self.stats["requests_sent"] += hp["F"]
        self.stats["request_bytes"] += 16 * hp["F"]
        self.stats["answer_bytes"] += 8 * k * hp["F"]

    def sample(self, inputs: Tensor, seed: Optional[int] = None, batch_base: int = 0) -> PartitionedBatches:
        """inputs [B, S] i64 (this rank's batches; device or pinned host) -> PartitionedBatches"""
        if inputs.dim() != 2 or inputs.dtype != torch.int64 or tuple(inputs.shape) != (self.B, self.S):
            raise ValueError(f"inputs must be an int64 tensor of shape {(self.B, self.S)}")
        seed = _rng_get() if seed is None else seed
        dev, S = self.device, self.S
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            self.samples[:, :S].copy_(inputs, non_blocking=True)
            self.err.zero_()
            self.lens.zero_()
            self.lens[0, 0].fill_(S)
            marks = [] if self.profile is not None else None

            def on(g):
                return torch.cuda.stream(g.stream) if g.stream is not None else torch.cuda.stream(main)

            for g in self.groups:
                if g.stream is not None:
                    g.stream.wait_stream(main)
                with on(g):
                    self._mark(g, marks, "start")
            for h, k in enumerate(self.fanouts):
                # phase-major issue order: while the host waits for one group's counts and that group's
                # all-to-all runs, the other group's kernels are already queued on its own stream
                for g in self.groups:
                    with on(g):
                        self._begin(g, h, batch_base)
                        self._mark(g, marks, "bucket")
                for g in self.groups:
                    with on(g):
                        self._requests(g)
                        self._mark(g, marks, "a2a_requests")
                        self._serve(g, k, seed)
                        self._mark(g, marks, "serve")
                for g in self.groups:
                    with on(g):
                        self._answers(g, k)
                        self._mark(g, marks, "a2a_answers")
                        self._finish(g, h, k, batch_base)
                        self._mark(g, marks, "layout")
            for g in self.groups:
                if g.stream is not None:
                    main.wait_stream(g.stream)
                g.hop = {}
            host = torch.cat([self.lens.reshape(-1), self.err.to(torch.int64)]).cpu().numpy()   # the call's last sync
        if marks is not None:
            for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
                self.profile[name] = self.profile.get(name, 0.0) + e0.elapsed_time(e1)
            self.profile["calls"] = self.profile.get("calls", 0) + 1
        N.check(N.lib.tchgeo_status_from_error_word(int(host[-1]) & 0xFFFFFFFF))
        lens = host[:-1].reshape(2, len(self.fanouts) + 1, self.B)
        return PartitionedBatches(self, lens[0], lens[1])


# ---------------------------------------------------------------------------------------------------------------
# Device-only protocol ("fixed segments", csrc/partitioned_fixed.cu): both exchanges of a hop are peer-memory stores
# into fixed per-pair segments, every count stays on the device, and the host never waits inside a step.
# ---------------------------------------------------------------------------------------------------------------
def segment_rows(num_batches: int, frontier_cap: int, world: int, slack: float) -> int:
    """rows a (requester, owner) pair owns in a hop whose frontier has at most num_batches * frontier_cap nodes"""
    total = num_batches * frontier_cap
    if world == 1:
        return max(total, 1) + (total & 1)
    rows = min(max(int(slack * total / world) + 1024, 1), max(total, 1))
    return rows + (rows & 1)      # even: a segment's answer rows then start 16-byte aligned (vector stores)


def frontier_caps(seeds_per_batch: int, fanouts: Sequence[int]) -> List[int]:
    """worst-case frontier size per batch of every hop (the recurrence of tchgeo_neighbor_sampling_capacity)"""
    caps, f = [], int(seeds_per_batch)
    for k in fanouts:
        caps.append(f)
        f *= int(k)
    return caps


class SegmentBuffers:
    """One rank's exchange buffers plus the device addresses of every rank's: req_in [world, seg, 2] i64 and cnt_in
    [world] i64 (written by the requesters), ans_in [world, seg, 2k] i32 (written by the owners).  `barrier()` orders
    the ranks on the current stream."""

    def __init__(self, req_in, cnt_in, ans_in, req_ptrs, cnt_ptrs, ans_ptrs, barrier):
        self.req_in, self.cnt_in, self.ans_in = req_in, cnt_in, ans_in
        self.req_ptrs = np.array([int(x) for x in req_ptrs], dtype=np.uint64)
        self.cnt_ptrs = np.array([int(x) for x in cnt_ptrs], dtype=np.uint64)
        self.ans_ptrs = np.array([int(x) for x in ans_ptrs], dtype=np.uint64)
        self.barrier = barrier

    @staticmethod
    def sizes(num_batches, capF, fanouts, world, slack):
        segs = [segment_rows(num_batches, f, world, slack) for f in capF]
        req_words = world * max(segs, default=1) * 2
        ans_words = world * max((s * 2 * max(k, 1) for s, k in zip(segs, fanouts)), default=1)
        return segs, req_words, ans_words

    @staticmethod
    def symmetric(num_batches, capF, fanouts, comm, device, slack):
        """torch symmetric memory over the communicator's group (NVLink peer memory): allocation, the rendezvous that
        gives every rank every peer's device address and the stream-ordered barrier are torch's plumbing; the stores
        are this library's kernels."""
        import torch.distributed._symmetric_memory as symm
        group = comm.group if comm.group is not None else dist.group.WORLD
        _, req_words, ans_words = SegmentBuffers.sizes(num_batches, capF, fanouts, comm.world, slack)
        with torch.cuda.device(device):
            req = symm.empty(req_words, dtype=torch.int64, device=device)
            cnt = symm.empty(max(comm.world, 16), dtype=torch.int64, device=device)
            ans = symm.empty(ans_words, dtype=torch.int32, device=device)
            hdls = [symm.rendezvous(t, group) for t in (req, cnt, ans)]
        for h in hdls:
            if len(h.buffer_ptrs) != comm.world or int(h.rank) != comm.rank:
                raise RuntimeError("symmetric memory rendezvous does not match the communicator")
        cnt.zero_()
        out = SegmentBuffers(req, cnt, ans, hdls[0].buffer_ptrs, hdls[1].buffer_ptrs, hdls[2].buffer_ptrs,
                             lambda: hdls[2].barrier(channel=0))
        out._hdls = hdls
        return out

    @staticmethod
    def virtual(num_batches, capF, fanouts, world, device, slack):
        """`world` ranks' buffers on ONE device (tests, single-GPU runs): plain tensors, stream order is the barrier.
        -> list of SegmentBuffers, one per virtual rank."""
        _, req_words, ans_words = SegmentBuffers.sizes(num_batches, capF, fanouts, world, slack)
        reqs = [torch.empty(req_words, dtype=torch.int64, device=device) for _ in range(world)]
        cnts = [torch.zeros(max(world, 16), dtype=torch.int64, device=device) for _ in range(world)]
        anss = [torch.empty(ans_words, dtype=torch.int32, device=device) for _ in range(world)]
        ptrs = lambda ts: [t.data_ptr() for t in ts]
        return [SegmentBuffers(reqs[r], cnts[r], anss[r], ptrs(reqs), ptrs(cnts), ptrs(anss), lambda: None)
                for r in range(world)]


class PartitionedPlanF:
    """neighbor_sampling_homogenous over a column-partitioned CSC with the device-only exchange.  `sample` is collective
    (every rank calls it the same number of times); the phase methods (`begin`, `scatter`, `serve`, `finish`, `end`)
    let a test drive several virtual ranks of one device in lock step."""

    def __init__(self, part: ColumnPartition, num_batches: int, seeds_per_batch: int, num_neighbors: Sequence[int],
                 sampler=None, world: Optional[int] = None, rank: Optional[int] = None, comm=None, buffers=None,
                 edge_bases=None, slack: float = 1.5, outputs=None, indices32=None):
        """outputs: optional (samples, rows, cols, edge_index) [B, capacity] tensors to write into (views of a larger
        plan's buffers); indices32: an int32 replica of part.indices that already exists."""
        self.part = part
        self.fanouts = [int(k) for k in num_neighbors]
        self.kind, _ = _extract_sampler(sampler, hetero=False)
        if self.kind == N.SAMPLER_WEIGHTED and part.weights is None:
            raise ValueError("weighted sampling needs ColumnPartition.weights_local")
        self.comm = comm
        self.world = int(world if world is not None else (comm.world if comm is not None else 1))
        self.rank = int(rank if rank is not None else (comm.rank if comm is not None else 0))
        _check(part.ptrs, torch.int64, "col_ptrs_local")
        dev = part.ptrs.device
        _check(part.indices, torch.int64, "row_indices_local", dev)
        if part.num_nodes >= 2 ** 31 or part.indices.numel() >= 2 ** 31:
            raise ValueError("the compact answer rows need node ids and a rank's CSC share below 2^31")
        self.device, self.B, self.S = dev, int(num_batches), int(seeds_per_batch)
        B, S, H = self.B, self.S, len(self.fanouts)
        self.capF, cap_e, f = [], 0, S
        for k in self.fanouts:
            self.capF.append(f)
            f *= k
            cap_e += f
        self.cap_n, self.cap_e = S + cap_e, max(cap_e, 1)
        self.slack = 1.0 if self.world == 1 else float(slack)
        self.segs, _, _ = SegmentBuffers.sizes(B, self.capF, self.fanouts, self.world, self.slack)
        if buffers is None:
            if self.world == 1:
                buffers = SegmentBuffers.virtual(B, self.capF, self.fanouts, 1, dev, self.slack)[0]
            else:
                buffers = SegmentBuffers.symmetric(B, self.capF, self.fanouts, comm, dev, self.slack)
        self.buf = buffers
        i64 = dict(dtype=torch.int64, device=dev)
        if outputs is not None:
            self.samples, self.rows, self.cols, self.eidx = outputs
            for t, c in ((self.samples, self.cap_n), (self.rows, self.cap_e), (self.cols, self.cap_e), (self.eidx, self.cap_e)):
                if tuple(t.shape) != (B, c) or not t.is_contiguous():
                    raise ValueError("outputs must be contiguous [num_batches, capacity] tensors")
        else:
            self.samples = torch.empty((B, self.cap_n), **i64)
            self.rows = torch.empty((B, self.cap_e), **i64)
            self.cols = torch.empty((B, self.cap_e), **i64)
            self.eidx = torch.empty((B, self.cap_e), **i64)
        self.lens = torch.zeros((2, H + 1, B), **i64)               # [0] node_len, [1] edge_len after h hops
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        fmax = max(self.capF, default=0)
        self.send = torch.empty(self.world * max(self.segs, default=1) * 2, **i64)
        self.cursor = torch.zeros(max(self.world, 16), **i64)
        self.slot_of = torch.empty(max(B * fmax, 1), dtype=torch.int32, device=dev)
        ws = max((N.lib.tchgeo_partf_workspace_bytes(B, c) for c in self.capF), default=0)
        self.ws = torch.empty(max(int(ws), 1), dtype=torch.uint8, device=dev)
        # int32 replica of this rank's share of row_indices: the serve kernel's random gathers span half the DRAM lines
        self.indices32 = indices32
        if indices32 is None and part.indices.numel() and os.environ.get("TCHGEO_INDEX_REPLICA", "1") != "0":
            i32 = torch.empty(part.indices.numel(), dtype=torch.int32, device=dev)
            scratch = torch.empty(1, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                st = N.lib.tchgeo_compress_indices(_ptr(part.indices), part.indices.numel(), _ptr(i32), _ptr(scratch), _stream(dev))
            if st == N.OK:
                self.indices32 = i32
            elif st != N.ERR_INDEX:
                N.check(st)
        if edge_bases is None:
            edge_bases = comm.all_gather_int(part.edge_base, dev) if comm is not None else [part.edge_base]
        if len(edge_bases) != self.world:
            raise ValueError("edge_bases must have one entry per rank")
        self.edge_bases = torch.tensor([int(x) for x in edge_bases], **i64)
        self.stats = {"requests_sent": 0, "request_bytes": 0, "answer_bytes": 0}
        self.profile = None
        self.peer = self.buf          # (bench.py reports which exchange ran)
        self.num_groups = 1

    # ---- phases (all asynchronous on the current stream) -------------------------------------------------
    def begin(self, inputs: Tensor, seed: int, batch_base: int):
        if inputs.dim() != 2 or inputs.dtype != torch.int64 or tuple(inputs.shape) != (self.B, self.S):
            raise ValueError(f"inputs must be an int64 tensor of shape {(self.B, self.S)}")
        self._seed, self._batch_base = seed, batch_base
        self.samples[:, :self.S].copy_(inputs, non_blocking=True)
        self.err.zero_()
        self.lens.zero_()
        self.lens[0, 0].fill_(self.S)

    def _frontier(self, h):
        lens = self.lens
        return (lens[0, h - 1] if h > 0 else None), lens[0, h]

    def scatter(self, h):
        fb, fe = self._frontier(h)
        with torch.cuda.device(self.device):
            N.check(N.lib.tchgeo_partf_scatter(_ptr(self.samples), self.cap_n, _ptr(fb), _ptr(fe), self.B, self.capF[h],
                                               self.part.cols_per_rank, self.world, self.rank, self._batch_base,
                                               self.segs[h], _ptr(self.send), _ptr(self.cursor), _ptr(self.slot_of),
                                               self.buf.req_ptrs.ctypes.data, self.buf.cnt_ptrs.ctypes.data,
                                               _ptr(self.err), _stream(self.device)))

    def serve(self, h):
        part, k = self.part, self.fanouts[h]
        with torch.cuda.device(self.device):
            N.check(N.lib.tchgeo_partf_serve(_ptr(part.ptrs), _ptr(part.indices), _ptr(self.indices32), _ptr(part.weights),
                                             part.col_begin, part.col_end - part.col_begin, part.indices.numel(),
                                             _ptr(self.buf.req_in), _ptr(self.buf.cnt_in), self.segs[h], 0, k, self.kind,
                                             self._seed, 0, self.world, self.rank, self.buf.ans_ptrs.ctypes.data,
                                             _ptr(self.err), _stream(self.device)))

    def finish(self, h):
        k, lens = self.fanouts[h], self.lens
        fb, fe = self._frontier(h)
        with torch.cuda.device(self.device):
            N.check(N.lib.tchgeo_partf_finish(_ptr(self.buf.ans_in), _ptr(self.slot_of), self.segs[h], k,
                                              _ptr(self.edge_bases), self.world, _ptr(fb), _ptr(fe), self.B, self.capF[h],
                                              _ptr(lens[0, h]), _ptr(lens[1, h]), _ptr(lens[0, h + 1]), _ptr(lens[1, h + 1]),
                                              _ptr(self.samples), self.cap_n, _ptr(self.rows), _ptr(self.cols),
                                              _ptr(self.eidx), self.cap_e, _ptr(self.err), _ptr(self.ws), self.ws.numel(),
                                              _stream(self.device)))

    def end(self) -> PartitionedBatches:
        host = torch.cat([self.lens.reshape(-1), self.err.to(torch.int64)]).cpu().numpy()   # the call's only sync
        N.check(N.lib.tchgeo_status_from_error_word(int(host[-1]) & 0xFFFFFFFF))
        lens = host[:-1].reshape(2, len(self.fanouts) + 1, self.B)
        front = lens[0, :-1].sum(axis=1) - np.concatenate([[0], lens[0, :-2].sum(axis=1)]) if len(self.fanouts) else []
        for F, k in zip(front, self.fanouts):
            self.stats["requests_sent"] += int(F)
            self.stats["request_bytes"] += 16 * int(F)
            self.stats["answer_bytes"] += 8 * k * int(F)
        return PartitionedBatches(self, lens[0], lens[1])

    def _mark(self, marks, name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    def sample(self, inputs: Tensor, seed: Optional[int] = None, batch_base: int = 0) -> PartitionedBatches:
        """inputs [B, S] i64 (this rank's batches; device or pinned host) -> PartitionedBatches"""
        seed = _rng_get() if seed is None else seed
        marks = [] if self.profile is not None else None
        with torch.cuda.device(self.device):
            self.begin(inputs, seed, batch_base)
            self._mark(marks, "start")
            for h in range(len(self.fanouts)):
                self.scatter(h)
                self._mark(marks, "scatter_put")
                self.buf.barrier()       # every requester's rows and counts are in every owner's buffers
                self._mark(marks, "barrier_requests")
                self.serve(h)
                self._mark(marks, "serve")
                self.buf.barrier()       # every owner's answer rows are in every requester's buffer
                self._mark(marks, "barrier_answers")
                self.finish(h)
                self._mark(marks, "layout")
            out = self.end()
        if marks is not None:
            for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
                self.profile[name] = self.profile.get(name, 0.0) + e0.elapsed_time(e1)
            self.profile["calls"] = self.profile.get("calls", 0) + 1
        return out


def sample_virtual_ranks(plans: List[PartitionedPlanF], inputs: List[Tensor], seed: int, batch_bases: List[int]):
    """Drive the plans of `world` virtual ranks that share ONE device through a step in lock step (phase by phase, in
    stream order): what the barriers do on real ranks.  -> one PartitionedBatches per rank."""
    for p, x, bb in zip(plans, inputs, batch_bases):
        p.begin(x, seed, bb)
    for h in range(len(plans[0].fanouts)):
        for p in plans:
            p.scatter(h)
        for p in plans:
            p.serve(h)
        for p in plans:
            p.finish(h)
    return [p.end() for p in plans]


class PartitionedPlanGroups:
    """PartitionedPlanF over `groups` contiguous ranges of a rank's batches, every range with its own exchange buffers
    and its own stream: while one group's rows are in flight over NVLink (and its barrier waits for the slowest rank),
    the other group's kernels run, so the exchange is hidden behind compute instead of being added to it.  Results land
    in one set of [B, capacity] buffers.  `sample` is collective."""

    def __init__(self, part: ColumnPartition, num_batches: int, seeds_per_batch: int, num_neighbors: Sequence[int],
                 sampler=None, comm=None, groups: int = 2, slack: float = 1.5, edge_bases=None):
        dev = part.ptrs.device
        self.B, self.S, self.device = int(num_batches), int(seeds_per_batch), dev
        self.fanouts = [int(k) for k in num_neighbors]
        G = max(1, min(int(groups), self.B))
        cuts = [self.B * g // G for g in range(G + 1)]
        caps = frontier_caps(self.S, self.fanouts)
        cap_e = sum(c * k for c, k in zip(caps, self.fanouts))
        self.cap_n, self.cap_e = self.S + cap_e, max(cap_e, 1)
        i64 = dict(dtype=torch.int64, device=dev)
        self.samples = torch.empty((self.B, self.cap_n), **i64)
        self.rows = torch.empty((self.B, self.cap_e), **i64)
        self.cols = torch.empty((self.B, self.cap_e), **i64)
        self.eidx = torch.empty((self.B, self.cap_e), **i64)
        if edge_bases is None and comm is not None:
            edge_bases = comm.all_gather_int(part.edge_base, dev)
        self.plans, i32 = [], None
        for g in range(G):
            b0, b1 = cuts[g], cuts[g + 1]
            outs = (self.samples[b0:b1], self.rows[b0:b1], self.cols[b0:b1], self.eidx[b0:b1])
            pl = PartitionedPlanF(part, b1 - b0, self.S, self.fanouts, sampler, comm=comm, slack=slack, outputs=outs,
                                  indices32=i32, edge_bases=edge_bases)
            i32 = pl.indices32
            pl.b0 = b0
            self.plans.append(pl)
        with torch.cuda.device(dev):
            self.streams = [torch.cuda.Stream(device=dev) for _ in self.plans]
        self.comm, self.world, self.rank = comm, self.plans[0].world, self.plans[0].rank
        self.num_groups = G
        self.peer = self.plans[0].buf
        self.profile = None
        self.slack = self.plans[0].slack

    @property
    def stats(self):
        return {k: sum(p.stats[k] for p in self.plans) for k in self.plans[0].stats}

    def sample(self, inputs: Tensor, seed: Optional[int] = None, batch_base: int = 0) -> PartitionedBatches:
        if inputs.dim() != 2 or inputs.dtype != torch.int64 or tuple(inputs.shape) != (self.B, self.S):
            raise ValueError(f"inputs must be an int64 tensor of shape {(self.B, self.S)}")
        seed = _rng_get() if seed is None else seed
        marks = [] if self.profile is not None else None
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            for pl, st in zip(self.plans, self.streams):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    pl.begin(inputs[pl.b0:pl.b0 + pl.B], seed, batch_base + pl.b0)
            # phase-major issue order: both groups' kernels of a phase are queued before anyone reaches a barrier
            for h in range(len(self.fanouts)):
                for pl, st in zip(self.plans, self.streams):
                    with torch.cuda.stream(st):
                        pl.scatter(h)
                for pl, st in zip(self.plans, self.streams):
                    with torch.cuda.stream(st):
                        pl.buf.barrier()
                        pl.serve(h)
                for pl, st in zip(self.plans, self.streams):
                    with torch.cuda.stream(st):
                        pl.buf.barrier()
                        pl.finish(h)
            outs = []
            for pl, st in zip(self.plans, self.streams):
                with torch.cuda.stream(st):
                    outs.append(pl.end())
                main.wait_stream(st)
        if marks is not None:
            self.profile["calls"] = self.profile.get("calls", 0) + 1
        node_len = np.concatenate([o._node_len for o in outs], axis=1)
        edge_len = np.concatenate([o._edge_len for o in outs], axis=1)
        return PartitionedBatches(self, node_len, edge_len)
