"""ctypes front-end of the CPU oracle (oracle/tchgeo_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  Nothing under tch-geometric_b200/ imports this module.

All functions take and return numpy int64/float64 arrays and mirror the reference's Python surface
(tch_geometric/tch_geometric.pyi) closely enough that tests read like the reference's tests.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtchgeo_oracle.so")

RNG_XOSHIRO = 0
RNG_COUNTER = 1
SAMPLER_UNIFORM = 0
SAMPLER_UNIFORM_REPLACE = 1
SAMPLER_WEIGHTED = 2
FILTER_NONE = -1
TEMPORAL_STATIC, TEMPORAL_RELATIVE, TEMPORAL_DYNAMIC = 0, 1, 2

OK, ERR_ARG, ERR_CAPACITY, ERR_PANIC = 0, 1, 2, 3


class OraclePanic(RuntimeError):
    """The reference would panic on this input (e.g. gen_range(0..0), out-of-range seed)."""


def build(force=False):
    src = os.path.join(_HERE, "tchgeo_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-C", _HERE], env=env, stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()  # no-op unless the library is missing or older than its source
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, ty=ctypes.c_int64):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.POINTER(ty))


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def _check(rc):
    if rc == ERR_PANIC:
        raise OraclePanic("reference would panic")
    if rc == ERR_CAPACITY:
        raise MemoryError("oracle output capacity exceeded")
    if rc != OK:
        raise ValueError(f"oracle error {rc}")


def philox4x32_10(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def kat_xoshiro(state, n):
    st = (ctypes.c_uint64 * 4)(*state)
    out = (ctypes.c_uint64 * n)()
    lib().orc_kat_xoshiro(st, ctypes.c_int64(n), out)
    return [int(x) for x in out]


def kat_seed_from_u64(seed):
    out = (ctypes.c_uint64 * 4)()
    lib().orc_kat_seed_from_u64(ctypes.c_uint64(seed), out)
    return [int(x) for x in out]


def kat_reduce(seed, kind, n, range_=1, high=1.0):
    """kind 0: gen_range(0..range_) values; 1: f32 bits of gen_range(0.0..1.0); 2: f64 bits of gen_range(0.0..high)"""
    out = (ctypes.c_uint64 * n)()
    lib().orc_kat_reduce(ctypes.c_uint64(seed), ctypes.c_int(kind), ctypes.c_uint64(range_), ctypes.c_double(high),
                         ctypes.c_int64(n), out)
    return [int(x) for x in out]


def ind2ptr(ind, m):
    ind = _i64(ind)
    out = np.empty(m + 1, dtype=np.int64)
    _check(lib().orc_ind2ptr(_p(ind), ctypes.c_int64(ind.size), ctypes.c_int64(m), _p(out)))
    return out


def _size(size):
    if isinstance(size, (tuple, list)):
        return int(size[0]), int(size[1])
    return int(size), int(size)


def _to_csx(row_col, size, csc):
    row_col = _i64(row_col)
    s0, s1 = _size(size)
    E = row_col.shape[1]
    row = np.ascontiguousarray(row_col[0])
    col = np.ascontiguousarray(row_col[1])
    ptrs = np.empty((s1 if csc else s0) + 1, dtype=np.int64)
    indices = np.empty(E, dtype=np.int64)
    perm = np.empty(E, dtype=np.int64)
    _check(lib().orc_to_csx(_p(row), _p(col), ctypes.c_int64(E), ctypes.c_int64(s0), ctypes.c_int64(s1),
                            ctypes.c_int(1 if csc else 0), _p(ptrs), _p(indices), _p(perm)))
    return ptrs, indices, perm


def to_csc(row_col, size):
    """python.rs:27-39"""
    return _to_csx(row_col, size, True)


def to_csr(row_col, size):
    """python.rs:41-53"""
    return _to_csx(row_col, size, False)


def csc_edge_cumsum(col_ptrs, row_data):
    """transform.rs:36-60 (returns the result instead of mutating in place)"""
    col_ptrs = _i64(col_ptrs)
    out = np.ascontiguousarray(np.asarray(row_data, dtype=np.float64)).copy()
    _check(lib().orc_csc_edge_cumsum_f64(_p(col_ptrs), ctypes.c_int64(col_ptrs.size - 1), _p(out, ctypes.c_double),
                                         ctypes.c_int64(out.size)))
    return out


def csc_sort_edges(col_ptrs, perm, row_weights, descending=False):
    """transform.rs:7-34: per column, new_perm[col] = perm[col][argsort(weights[col], descending)].  torch's argsort
    is unstable (ties unspecified); ties are resolved by CSC position here, as on the GPU."""
    col_ptrs, perm = _i64(col_ptrs), _i64(perm)
    w = np.asarray(row_weights, dtype=np.float64)
    out = perm.copy()  # :15
    for s, e in zip(col_ptrs[:-1], col_ptrs[1:]):
        if e - s <= 1:  # :22-24
            continue
        e = min(int(e), perm.size)  # Tensor::slice clamps
        keys = -w[s:e] if descending else w[s:e]
        out[s:e] = perm[s:e][np.argsort(keys, kind="stable")]
    return out


def has_edge(ptrs, indices, x, y):
    return bool(lib().orc_has_edge(_p(_i64(ptrs)), _p(_i64(indices)), ctypes.c_int64(x), ctypes.c_int64(y)))


def _sampler_args(sampler):
    """sampler: None | ('uniform', with_replacement) | ('weighted', weights)"""
    if sampler is None:
        return SAMPLER_UNIFORM, None
    kind, arg = sampler
    if kind == "uniform":
        return (SAMPLER_UNIFORM_REPLACE if arg else SAMPLER_UNIFORM), None
    if kind == "weighted":
        return SAMPLER_WEIGHTED, arg
    raise ValueError(kind)


def _worst_case(num_inputs, fanouts):
    nodes, edges, f = num_inputs, 0, num_inputs
    for k in fanouts:
        f = f * k
        edges += f
        nodes += f
    return nodes, edges


def neighbor_sampling_homogenous(col_ptrs, row_indices, inputs, num_neighbors, sampler=None, filter=None,
                                 rng_mode=RNG_COUNTER, seed=0, batch=0):
    """python.rs:187-271 -> (samples, rows, cols, edge_index, layer_offsets)

    filter: None | dict(mode, forward, window=(lo, hi), timestamps, inputs_state)
    """
    col_ptrs, row_indices, inputs = _i64(col_ptrs), _i64(row_indices), _i64(inputs)
    fan = _i64(num_neighbors)
    kind, w = _sampler_args(sampler)
    w = None if w is None else np.ascontiguousarray(np.asarray(w, dtype=np.float64))
    cap_n, cap_e = _worst_case(inputs.size, [int(k) for k in fan])
    samples = np.empty(max(cap_n, 1), dtype=np.int64)
    rows = np.empty(max(cap_e, 1), dtype=np.int64)
    cols = np.empty(max(cap_e, 1), dtype=np.int64)
    eidx = np.empty(max(cap_e, 1), dtype=np.int64)
    lens = np.zeros(2, dtype=np.int64)
    lo = np.zeros(3 * max(fan.size, 1), dtype=np.int64)
    if filter is None:
        fm, ff, wl, wh, ts, st = FILTER_NONE, 0, 0, 0, None, None
    else:
        fm, ff = int(filter["mode"]), int(bool(filter["forward"]))
        wl, wh = (int(x) for x in filter["window"])
        ts, st = _i64(filter["timestamps"]), _i64(filter["inputs_state"])
    rc = lib().orc_neighbor_sampling_homogenous(
        _p(col_ptrs), ctypes.c_int64(col_ptrs.size - 1), _p(row_indices),
        _p(inputs), ctypes.c_int64(inputs.size), _p(fan), ctypes.c_int(fan.size),
        ctypes.c_int(kind), _p(w, ctypes.c_double),
        ctypes.c_int(fm), ctypes.c_int(ff), ctypes.c_int64(wl), ctypes.c_int64(wh), _p(ts), _p(st),
        ctypes.c_int(rng_mode), ctypes.c_uint64(seed), ctypes.c_uint32(batch),
        _p(samples), ctypes.c_int64(cap_n), _p(rows), _p(cols), _p(eidx), ctypes.c_int64(cap_e),
        _p(lens), _p(lo))
    _check(rc)
    ns, ne = int(lens[0]), int(lens[1])
    layer_offsets = [tuple(int(x) for x in lo[3 * h:3 * h + 3]) for h in range(fan.size)]
    return samples[:ns].copy(), rows[:ne].copy(), cols[:ne].copy(), eidx[:ne].copy(), layer_offsets


def neighbor_sampling_homogenous_batches(col_ptrs, row_indices, inputs, num_neighbors, sampler=None,
                                         rng_mode=RNG_XOSHIRO, seed=0, num_threads=0):
    """CPU-baseline driver: inputs [B, S]; returns (total_samples, total_edges)."""
    col_ptrs, row_indices, inputs = _i64(col_ptrs), _i64(row_indices), _i64(inputs)
    fan = _i64(num_neighbors)
    kind, w = _sampler_args(sampler)
    w = None if w is None else np.ascontiguousarray(np.asarray(w, dtype=np.float64))
    ts, te = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = lib().orc_neighbor_sampling_homogenous_batches(
        _p(col_ptrs), ctypes.c_int64(col_ptrs.size - 1), _p(row_indices), _p(inputs),
        ctypes.c_int64(inputs.shape[0]), ctypes.c_int64(inputs.shape[1]), _p(fan), ctypes.c_int(fan.size),
        ctypes.c_int(kind), _p(w, ctypes.c_double), ctypes.c_int(rng_mode), ctypes.c_uint64(seed),
        ctypes.c_int(num_threads), ctypes.byref(ts), ctypes.byref(te))
    _check(rc)
    return ts.value, te.value


def rel_key(edge_type):
    """neighbor_sampling.rs:255-258"""
    return "{}__{}__{}".format(*edge_type)


def neighbor_sampling_heterogenous(node_types, edge_types, col_ptrs, row_indices, inputs, num_neighbors, num_hops,
                                   sampler=None, filter=None, rng_mode=RNG_COUNTER, seed=0, batch=0):
    """python.rs:273-395 -> (samples{type}, rows{rel}, cols{rel}, edge_index{rel}, layer_offsets{rel})

    Relations are visited in `edge_types` order (canonicalisation of quirk Q6).
    sampler: None | ('uniform', bool) | ('weighted', {rel: weights})
    """
    T, R = len(node_types), len(edge_types)
    tix = {t: i for i, t in enumerate(node_types)}
    rels = [rel_key(e) for e in edge_types]
    rel_src = np.array([tix[e[0]] for e in edge_types], dtype=np.int32)
    rel_dst = np.array([tix[e[2]] for e in edge_types], dtype=np.int32)
    kind, wdict = _sampler_args(sampler)
    keep = []  # keep arrays alive

    def parr(arrs, ty=ctypes.c_int64):
        out = (ctypes.POINTER(ty) * len(arrs))()
        for i, a in enumerate(arrs):
            out[i] = _p(a, ty) if a is not None else None
        return out

    cp = [_i64(col_ptrs[r]) if r in col_ptrs else np.zeros(1, dtype=np.int64) for r in rels]
    ri = [_i64(row_indices[r]) if r in row_indices else np.zeros(0, dtype=np.int64) for r in rels]
    ncols = np.array([a.size - 1 for a in cp], dtype=np.int64)
    inp = [_i64(inputs[t]) if t in inputs else np.zeros(0, dtype=np.int64) for t in node_types]
    ninp = np.array([a.size for a in inp], dtype=np.int64)
    active = np.array([1 if r in num_neighbors else 0 for r in rels], dtype=np.uint8)
    fan = np.zeros((R, max(num_hops, 1)), dtype=np.int64)
    for i, r in enumerate(rels):
        if r in num_neighbors:
            fan[i, :num_hops] = np.asarray(num_neighbors[r][:num_hops], dtype=np.int64)
    ws = None
    if kind == SAMPLER_WEIGHTED:
        ws = [np.ascontiguousarray(np.asarray(wdict[r], dtype=np.float64)) if r in wdict else None for r in rels]
    # worst-case capacities (same recurrence the host driver uses)
    front = ninp.astype(object).copy()
    ncap = ninp.astype(object).copy()
    ecap = np.zeros(R, dtype=object)
    for ell in range(num_hops):
        add = np.zeros(T, dtype=object)
        for i in range(R):
            if active[i]:
                n = int(front[rel_dst[i]]) * int(fan[i, ell])
                ecap[i] += n
                add[rel_src[i]] += n
        front = add
        ncap = ncap + add
    samples = [np.empty(max(int(c), 1), dtype=np.int64) for c in ncap]
    rows = [np.empty(max(int(c), 1), dtype=np.int64) for c in ecap]
    cols = [np.empty(max(int(c), 1), dtype=np.int64) for c in ecap]
    eidx = [np.empty(max(int(c), 1), dtype=np.int64) for c in ecap]
    ncap64 = np.array([int(c) for c in ncap], dtype=np.int64)
    ecap64 = np.array([int(c) for c in ecap], dtype=np.int64)
    slen = np.zeros(T, dtype=np.int64)
    elen = np.zeros(R, dtype=np.int64)
    lo = np.zeros(R * max(num_hops, 1) * 3, dtype=np.int64)
    lolen = np.zeros(R, dtype=np.int64)
    if filter is None:
        fm, ff, wl, wh, ts, st = FILTER_NONE, 0, 0, 0, None, None
    else:
        fm, ff = int(filter["mode"]), int(bool(filter["forward"]))
        wl, wh = (int(x) for x in filter["window"])
        ts = [_i64(filter["timestamps"][r]) if r in filter["timestamps"] else None for r in rels]
        st = [_i64(filter["inputs_state"][t]) if t in filter["inputs_state"] else None for t in node_types]
    keep += [cp, ri, inp, ws, ts, st]
    rc = lib().orc_neighbor_sampling_heterogenous(
        ctypes.c_int(T), ctypes.c_int(R), _p(rel_src, ctypes.c_int32), _p(rel_dst, ctypes.c_int32),
        parr(cp), _p(ncols), parr(ri), parr(inp), _p(ninp), _p(np.ascontiguousarray(fan)), _p(active, ctypes.c_uint8),
        ctypes.c_int(num_hops), ctypes.c_int(kind), parr(ws, ctypes.c_double) if ws is not None else None,
        ctypes.c_int(fm), ctypes.c_int(ff), ctypes.c_int64(wl), ctypes.c_int64(wh),
        parr(ts) if ts is not None else None, parr(st) if st is not None else None,
        ctypes.c_int(rng_mode), ctypes.c_uint64(seed), ctypes.c_uint32(batch),
        parr(samples), _p(ncap64), _p(slen), parr(rows), parr(cols), parr(eidx), _p(ecap64), _p(elen),
        _p(lo), _p(lolen))
    _check(rc)
    lo = lo.reshape(R, max(num_hops, 1), 3)
    out_s = {t: samples[i][:slen[i]].copy() for i, t in enumerate(node_types)}
    out_r = {r: rows[i][:elen[i]].copy() for i, r in enumerate(rels) if r in col_ptrs}
    out_c = {r: cols[i][:elen[i]].copy() for i, r in enumerate(rels) if r in col_ptrs}
    out_e = {r: eidx[i][:elen[i]].copy() for i, r in enumerate(rels) if r in col_ptrs}
    out_lo = {r: [tuple(int(x) for x in lo[i, h]) for h in range(int(lolen[i]))] for i, r in enumerate(rels)
              if r in col_ptrs}
    return out_s, out_r, out_c, out_e, out_lo


def random_walk(row_ptrs, col_indices, start, walk_length, p, q, rng_mode=RNG_COUNTER, seed=0, walker_base=0,
                return_attempts=False):
    """python.rs:583-608 -> walks [S, walk_length+1]"""
    row_ptrs, col_indices, start = _i64(row_ptrs), _i64(col_indices), _i64(start)
    walks = np.empty((start.size, walk_length + 1), dtype=np.int64)
    att = ctypes.c_int64(0)
    rc = lib().orc_random_walk(_p(row_ptrs), ctypes.c_int64(row_ptrs.size - 1), _p(col_indices), _p(start),
                               ctypes.c_int64(start.size), ctypes.c_int64(walk_length), ctypes.c_float(p),
                               ctypes.c_float(q), ctypes.c_int(rng_mode), ctypes.c_uint64(seed),
                               ctypes.c_int64(walker_base), _p(walks), ctypes.byref(att))
    _check(rc)
    return (walks, att.value) if return_attempts else walks


def random_walk_mt(row_ptrs, col_indices, start, walk_length, p, q, rng_mode=RNG_XOSHIRO, seed=0, num_threads=0):
    row_ptrs, col_indices, start = _i64(row_ptrs), _i64(col_indices), _i64(start)
    walks = np.empty((start.size, walk_length + 1), dtype=np.int64)
    att = ctypes.c_int64(0)
    rc = lib().orc_random_walk_mt(_p(row_ptrs), ctypes.c_int64(row_ptrs.size - 1), _p(col_indices), _p(start),
                                  ctypes.c_int64(start.size), ctypes.c_int64(walk_length), ctypes.c_float(p),
                                  ctypes.c_float(q), ctypes.c_int(rng_mode), ctypes.c_uint64(seed),
                                  ctypes.c_int(num_threads), _p(walks), ctypes.byref(att))
    _check(rc)
    return walks, att.value


def _negative(node_types, edge_types, row_ptrs, col_indices, sizes, inputs, num_neg, try_count, inbound, heterogenous,
              rng_mode, seed):
    T, R = len(node_types), len(edge_types)
    tix = {t: i for i, t in enumerate(node_types)}
    rels = [rel_key(e) for e in edge_types]
    rel_src = np.array([tix[e[0]] for e in edge_types], dtype=np.int32)
    rel_dst = np.array([tix[e[2]] for e in edge_types], dtype=np.int32)

    def parr(arrs):
        out = (ctypes.POINTER(ctypes.c_int64) * len(arrs))()
        for i, a in enumerate(arrs):
            out[i] = _p(a)
        return out

    rp = [_i64(row_ptrs[r]) for r in rels]
    ci = [_i64(col_indices[r]) for r in rels]
    nrows = np.array([a.size - 1 for a in rp], dtype=np.int64)
    ncount = np.array([int(sizes[r][1]) for r in rels], dtype=np.int64)
    inp = [_i64(inputs[t]) if t in inputs else np.zeros(0, dtype=np.int64) for t in node_types]
    ninp = np.array([a.size for a in inp], dtype=np.int64)
    total = int(ninp.sum()) * int(num_neg)
    samples = [np.empty(max(int(ninp[t]) + total, 1), dtype=np.int64) for t in range(T)]
    rows = [np.empty(max(int(ninp[rel_src[r]]) * int(num_neg), 1), dtype=np.int64) for r in range(R)]
    cols = [np.empty_like(a) for a in rows]
    slen, elen = np.zeros(T, dtype=np.int64), np.zeros(R, dtype=np.int64)
    rc = lib().orc_negative_sampling(
        ctypes.c_int(T), ctypes.c_int(R), _p(rel_src, ctypes.c_int32), _p(rel_dst, ctypes.c_int32), parr(rp), parr(ci),
        _p(nrows), _p(ncount), parr(inp), _p(ninp), ctypes.c_int64(num_neg), ctypes.c_int64(try_count),
        ctypes.c_int(1 if inbound else 0), ctypes.c_int(1 if heterogenous else 0), ctypes.c_int(rng_mode),
        ctypes.c_uint64(seed), parr(samples), parr(rows), parr(cols), _p(slen), _p(elen))
    _check(rc)
    out_s = {t: samples[i][:slen[i]].copy() for i, t in enumerate(node_types)}
    out_r = {r: rows[i][:elen[i]].copy() for i, r in enumerate(rels)}
    out_c = {r: cols[i][:elen[i]].copy() for i, r in enumerate(rels)}
    counts = {t: int(ninp[i]) for i, t in enumerate(node_types)}
    return out_s, out_r, out_c, counts


def negative_sample_neighbors_homogenous(row_ptrs, col_indices, graph_size, inputs, num_neg, try_count,
                                         rng_mode=RNG_COUNTER, seed=0):
    """python.rs:689-721 -> (samples, rows, cols, sample_count)"""
    s, r, c, n = _negative(["n"], [("n", "e", "n")], {"n__e__n": row_ptrs}, {"n__e__n": col_indices},
                           {"n__e__n": _size(graph_size)}, {"n": inputs}, num_neg, try_count, False, False, rng_mode, seed)
    return s["n"], r["n__e__n"], c["n__e__n"], n["n"]


def negative_sample_neighbors_heterogenous(node_types, edge_types, row_ptrs, col_indices, sizes, inputs, num_neg,
                                           try_count, inbound, rng_mode=RNG_COUNTER, seed=0):
    """python.rs:723-783 -> (samples{type}, rows{rel}, cols{rel}, sample_count{type}); node types are visited in
    `node_types` order and relations in `edge_types` order (the reference iterates HashMaps)."""
    return _negative(node_types, edge_types, row_ptrs, col_indices, sizes, inputs, num_neg, try_count, inbound, True,
                     rng_mode, seed)


def tempo_random_walk(row_ptrs, col_indices, node_timestamps, edge_timestamps, start, start_timestamps, walk_length, window,
                      rng_mode=RNG_COUNTER, seed=0, walker_base=0):
    """python.rs:610-643 -> (walks, walk_timestamps), both [S, walk_length]"""
    row_ptrs, col_indices, start = _i64(row_ptrs), _i64(col_indices), _i64(start)
    nts, ets, sts = _i64(node_timestamps), _i64(edge_timestamps), _i64(start_timestamps)
    S, L = start.size, int(walk_length)
    walks = np.empty((S, max(L, 0)), dtype=np.int64)
    wts = np.empty_like(walks)
    rc = lib().orc_tempo_random_walk(_p(row_ptrs), ctypes.c_int64(row_ptrs.size - 1), _p(col_indices), _p(nts),
                                     ctypes.c_int64(nts.size), _p(ets), _p(start), _p(sts), ctypes.c_int64(S),
                                     ctypes.c_int64(L), ctypes.c_int64(int(window[0])), ctypes.c_int64(int(window[1])),
                                     ctypes.c_int(rng_mode), ctypes.c_uint64(seed), ctypes.c_int64(walker_base), _p(walks),
                                     _p(wts))
    _check(rc)
    return walks, wts


def unique_relabel(samples, num_seeds):
    """dedup stage (negative_sampling.rs:20-47 semantic) -> (nodes, local)"""
    samples = _i64(samples)
    nodes = np.empty(max(samples.size, 1), dtype=np.int64)
    local = np.empty(max(samples.size, 1), dtype=np.int64)
    n = ctypes.c_int64(0)
    _check(lib().orc_unique_relabel(_p(samples), ctypes.c_int64(samples.size), ctypes.c_int64(num_seeds),
                                    _p(nodes), ctypes.byref(n), _p(local)))
    return nodes[:n.value].copy(), local[:samples.size].copy()


def num_threads():
    return int(lib().orc_num_threads())


def serve_requests(ptrs_local, indices_local, col_begin, edge_base, req_ids, req_meta, fanout, sampler=None, seed=0, rel=0):
    """checker for tchgeo_serve_requests -> (out_ids [n, fanout], out_ptrs [n, fanout])"""
    ptrs_local, indices_local = _i64(ptrs_local), _i64(indices_local)
    req_ids, req_meta = _i64(req_ids), _i64(req_meta)
    kind, w = _sampler_args(sampler)
    w = None if w is None else np.ascontiguousarray(np.asarray(w, dtype=np.float64))
    n = req_ids.size
    out_ids = np.empty((n, fanout), dtype=np.int64)
    out_ptrs = np.empty((n, fanout), dtype=np.int64)
    _check(lib().orc_serve_requests(_p(ptrs_local), _p(indices_local), _p(w, ctypes.c_double), ctypes.c_int64(col_begin),
                                    ctypes.c_int64(ptrs_local.size - 1), ctypes.c_int64(edge_base), _p(req_ids),
                                    _p(req_meta), ctypes.c_int64(n), ctypes.c_int64(fanout), ctypes.c_int(kind),
                                    ctypes.c_uint64(seed), ctypes.c_uint32(rel), _p(out_ids), _p(out_ptrs)))
    return out_ids, out_ptrs
