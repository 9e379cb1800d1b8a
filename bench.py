#!/usr/bin/env python
"""bench.py -- headline benchmark of the sampling hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): synthetic ogbn-products-shaped graph (2 449 029 nodes,
61 859 140 CSC entries), homogeneous neighbor sampling, fanouts [15,10,5], 1024 seeds per batch,
default sampler (uniform without replacement).  One *step* = one pass of the hot path over
`--batches` (default 256) seed batches, i.e. the "many batches in flight" regime of SURVEY §8(d).

Prints ONE JSON line:
  value      sampled edges/s, whole job, seeds already resident in HBM when the timed region starts
             (CUDA events on the launching stream, max over ranks).
  e2e        the same metric through the public plan API with HOST buffers: seeds start in pinned host
             memory, every step's four output tensors + layer offsets are copied back to pinned host
             memory inside the timed region.
  roofline   hop-3 kernel (the dominant launch): algorithmic bytes 24*F + 40*E over its CUDA-event
             duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the CPU oracle (C restatement of the reference algorithm) on this box's host cores,
             on a bounded sample of the same workload.
  with_relabel  the same step with the dedup + insertion-order relabel stage (K7) enqueued after the hops.
  walk_steps_per_sec / hetero_edges_per_sec (+ "walk", "hetero" objects with their own roofline): BASELINE.json
             configs[2] (node2vec walk, walkers sharded over the ranks) and configs[3] (mag-shaped heterogeneous
             sampling, seed batches sharded), measured in the same run (--headline-only skips them).

--impl reference times the reference's CPU algorithm (the oracle port: the reference is Rust and
cannot be built in this image) with all host threads on bounded samples of the same workload.
Multi-GPU (torchrun, one rank per GPU): CSC replicated, seed batches sharded, no data-path
collective; weak scaling (each rank samples `--batches` batches per step).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tch-geometric_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from tools import synth  # noqa: E402

FANOUTS = [15, 10, 5]
SEEDS_PER_BATCH = 1024
METRIC = "sampled_edges_per_sec_3hop_15_10_5"
UNIT = "edges/s"
SAMPLER_LABEL = {"uniform": "uniform without replacement", "replace": "uniform with replacement",
                 "weighted": "weighted (f64 weights ~ U(0.2, 5.0) per CSC entry)"}


def metric_name(sampler):
    return METRIC if sampler == "uniform" else f"{METRIC}_{sampler}"


def edge_weights(num_edges, device):
    """w ~ U(0.2, 5.0) f64 per CSC position, mirroring the reference's own weighted test (neighbor_sampling.rs:475)."""
    g = torch.Generator(device=device)
    g.manual_seed(4242)
    return torch.rand(num_edges, generator=g, dtype=torch.float64, device=device) * 4.8 + 0.2


# Only the final JSON line may reach stdout: libraries (e.g. NCCL's version banner) print there too, so
# fd 1 is pointed at stderr for the whole run and the JSON goes to the saved descriptor.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def setup_device():
    """-> (rank, world, local, device); one NCCL process group per run when launched by torchrun"""
    import torch.distributed as dist
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    return rank, world, local, device


def build_graph(device, scale):
    t0 = time.time()
    ei, n = synth.products_like(device, scale=scale)
    if device.type == "cuda":
        torch.cuda.synchronize(device)
    log(f"[bench] graph generated: N={n} E={ei.shape[1]} in {time.time() - t0:.1f}s")
    return ei, n


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm
# ---------------------------------------------------------------------------------------------
def cpu_sampling_rate(ptrs, idx, n, num_batches, threads, first_batch, sampler=None):
    """sampler: None | ("uniform", True) | ("weighted", weights) in the oracle's notation"""
    from oracle import oracle as O
    seeds = synth.seed_batches(n, num_batches, SEEDS_PER_BATCH, first_batch=first_batch)
    t0 = time.perf_counter()
    ts, te = O.neighbor_sampling_homogenous_batches(ptrs, idx, seeds, FANOUTS, sampler=sampler, rng_mode=O.RNG_XOSHIRO,
                                                    seed=first_batch, num_threads=threads)
    dt = time.perf_counter() - t0
    return te / dt, te, dt


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation (oracle port) on the host cores."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    ei, n = build_graph(device, args.scale)
    ei = ei.cpu().numpy()
    t0 = time.time()
    ptrs, idx, _ = O.to_csc(ei, n)  # the reference's own path restated (storage.rs:103-127)
    log(f"[bench] reference arm: CSC built on the CPU in {time.time() - t0:.1f}s; {cores} host threads")
    del ei
    osampler = oracle_sampler(args.sampler, idx.size, device)
    per_step = args.batches if args.ref_batches <= 0 else args.ref_batches  # same batches per step as our arm
    for w in range(args.warmup):
        cpu_sampling_rate(ptrs, idx, n, per_step, cores, first_batch=w * per_step, sampler=osampler)
    tot_e, tot_t = 0, 0.0
    for k in range(args.steps):
        _, te, dt = cpu_sampling_rate(ptrs, idx, n, per_step, cores, first_batch=(args.warmup + k) * per_step,
                                      sampler=osampler)
        tot_e += te
        tot_t += dt
    value = tot_e / tot_t
    sample = f"{per_step} batches x {SEEDS_PER_BATCH} seeds per step, {args.steps} steps, {cores} threads"
    out = {
        "impl": "reference", "metric": metric_name(args.sampler), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(n, int(idx.size), per_step, args.scale, args.sampler),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def oracle_sampler(name, num_edges, device):
    if name == "replace":
        return ("uniform", True)
    if name == "weighted":
        return ("weighted", edge_weights(num_edges, device).cpu().numpy())
    return None


def workload_config(n, e, batches, scale, sampler="uniform"):
    return {
        "workload": f"products-shaped synthetic graph (N={n}, E={e}{'' if scale == 1.0 else f', scale={scale}'}), "
                    f"neighbor_sampling_homogenous fanouts {FANOUTS}, {SEEDS_PER_BATCH} seeds/batch, "
                    f"{batches} batches/step, {SAMPLER_LABEL[sampler]}",
        "fanouts": FANOUTS, "seeds_per_batch": SEEDS_PER_BATCH, "batches_per_step": batches,
        "l2_policy": "inputs larger than L2: row_indices is 495 MB and each step streams GBs of outputs through the "
                     "126 MB L2; seeds differ every step",
        "parallelism": "seed batches sharded over ranks, CSC replicated, no data-path collective",
    }


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import tch_geometric as thg
    import torch.distributed as dist

    rank, world, local, device = setup_device()
    B, S, K, W = args.batches, SEEDS_PER_BATCH, args.steps, args.warmup
    l2_fetch = None
    if args.l2_fetch:
        l2_fetch = thg.set_l2_fetch_granularity(args.l2_fetch, device)
        log(f"[bench] L2 fetch granularity set to {args.l2_fetch}, driver reports {l2_fetch}")

    # ---- setup (untimed): graph, CSC through the new to_csc -----------------------------------
    ei, n = build_graph(device, args.scale)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    to_csc_ms = []
    for _ in range(3):  # first call pays module load + workspace cudaMalloc; report the steady state too
        ptrs = idx = perm = None
        ev0.record()
        ptrs, idx, perm = thg.to_csc(ei, n)
        ev1.record()
        torch.cuda.synchronize()
        to_csc_ms.append(ev0.elapsed_time(ev1))
    to_csc_first_ms, to_csc_ms = to_csc_ms[0], min(to_csc_ms[1:])
    E = int(idx.numel())
    del ei, perm
    torch.cuda.empty_cache()
    log(f"[bench] rank {rank}: to_csc {to_csc_ms:.2f} ms (first call {to_csc_first_ms:.2f} ms) for E={E}")

    # every rank samples its own global batch indices: step s, rank r -> batches [(s*world + r)*B, +B)
    steps_total = W + 2 * K  # [W, W+K): serial pass with per-launch events; [W+K, W+2K): pipelined pass (value)
    host_seeds = torch.empty((steps_total, B, S), dtype=torch.int64).pin_memory()
    for s in range(steps_total):
        host_seeds[s] = torch.from_numpy(synth.seed_batches(n, B, S, first_batch=(s * world + rank) * B))
    dev_seeds = host_seeds.to(device)
    sampler, weights = None, None
    if args.sampler == "replace":
        sampler = thg.UniformEdgeSampler(with_replacement=True)
    elif args.sampler == "weighted":
        weights = edge_weights(E, device)
        sampler = thg.WeightedEdgeSampler(weights)
    plan = thg.HomogenousSampler(ptrs, idx, B, S, FANOUTS, sampler=sampler)
    cap_n, cap_e = int(plan._call.cap_n[0]), int(plan._call.cap_e[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) + per-launch roofline ---------------------------------
    for s in range(W):
        plan.sample(dev_seeds[s], seed=1000 + s, batch_base=(s * world + rank) * B, timed=True)
    clocks = ClockSampler(local)
    clocks.start()  # sampled over both timed regions (device-resident loop and e2e loop)
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    edges = 0
    hop_ms = np.zeros(len(FANOUTS))
    hop_F = np.zeros(len(FANOUTS))
    hop_E = np.zeros(len(FANOUTS))
    hop_deg = np.zeros(len(FANOUTS))   # sum of the frontier nodes' degrees (weighted: the weights that are scanned)
    start.record()
    for s in range(W, W + K):
        res = plan.sample(dev_seeds[s], seed=1000 + s, batch_base=(s * world + rank) * B, timed=True)
        edges += int(res.edges_len.sum())
        lo = res.layer_offsets  # [B, H, 3]
        n_end = np.concatenate([lo[:, 1:, 0], res.samples_len[:, None]], axis=1)  # len(samples) after each hop
        e_end = np.concatenate([lo[:, 1:, 1], res.edges_len[:, None]], axis=1)
        hop_E += (e_end - lo[:, :, 1]).sum(axis=0)
        f_begin = np.concatenate([np.zeros((B, 1), dtype=np.int64), lo[:, :-1, 0]], axis=1)
        hop_F += (lo[:, :, 0] - f_begin).sum(axis=0)
        hop_ms += res.launch_ms
    stop.record()
    if weights is not None:  # untimed: frontier degree sums of the last step, for the weighted byte model
        pos = torch.arange(cap_n, device=device)[None, :]
        f_b = torch.as_tensor(f_begin, device=device)
        f_e = torch.as_tensor(lo[:, :, 0], device=device)
        for h in range(len(FANOUTS)):
            ids = res.samples[(pos >= f_b[:, h:h + 1]) & (pos < f_e[:, h:h + 1])]
            hop_deg[h] = float((ptrs[ids + 1] - ptrs[ids]).sum().item()) * K
        del pos, ids
    barrier()
    serial_ms = start.elapsed_time(stop)
    from tch_geometric.sharding import reduce_job
    serial_ms, serial_edges_all = reduce_job(serial_ms, float(edges), device)
    serial = {"ms_per_step": serial_ms / K, "value": serial_edges_all / (serial_ms * 1e-3),
              "what": "one plan, one stream, host waits for every step's lengths before enqueueing the next step"}

    # ---- value: the same K-step job with two plans on two streams (HomogenousSampler.sample_async / result) ------
    # The host enqueues step s+1 before it waits for the lengths of step s, so the device never idles between steps.
    plans = [plan, thg.HomogenousSampler(ptrs, idx, B, S, FANOUTS, sampler=sampler)]
    streams = [torch.cuda.Stream(device), torch.cuda.Stream(device)]

    def pipelined(first, count):
        total, pending = 0, []
        for i, s in enumerate(range(first, first + count)):
            j = i & 1
            if len(pending) == 2:
                total += int(plans[pending.pop(0)].result().edges_len.sum())
            with torch.cuda.stream(streams[j]):
                plans[j].sample_async(dev_seeds[s], seed=1000 + s, batch_base=(s * world + rank) * B)
            pending.append(j)
        while pending:
            total += int(plans[pending.pop(0)].result().edges_len.sum())
        return total

    pipelined(0, max(W, 2))  # warm-up of both plans
    barrier()
    cur = torch.cuda.current_stream(device)
    for st in streams:
        st.wait_stream(cur)
    start.record(cur)
    for st in streams:
        st.wait_stream(cur)
    t0 = time.perf_counter()
    edges = pipelined(W + K, K)
    for st in streams:
        cur.wait_stream(st)
    stop.record(cur)
    barrier()
    elapsed_ms = max(start.elapsed_time(stop), 0.0)
    wall_ms = (time.perf_counter() - t0) * 1e3
    elapsed_ms, edges_all = reduce_job(elapsed_ms, float(edges), device)
    value = edges_all / (elapsed_ms * 1e-3)
    seeds_per_s = world * B * S * K / (elapsed_ms * 1e-3)
    serial["pipelined_wall_ms_per_step"] = wall_ms / K

    # kernels of ours per step: fill_i64_kernel + per hop (frontier_max_kernel +) the hop kernel
    hop_kernel_name = "hop_kernel"
    launches_per_step = plan.num_launches  # fill_i64_kernel + one hop kernel per hop (counted by the library)
    peak, peak_src = measured_peak_gbs()
    dom = int(np.argmax(hop_ms))
    # per launch, summed over the K timed launches; the weighted sampler adds 8 B per scanned weight
    alg_bytes = 24.0 * hop_F + 40.0 * hop_E + 8.0 * hop_deg
    achieved = alg_bytes[dom] / (hop_ms[dom] * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": f"{hop_kernel_name}<{args.sampler.upper()}> hop {dom + 1} (fanout {FANOUTS[dom]})",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes[dom] / K, "launch_ms": hop_ms[dom] / K,
        "per_hop": [{"hop": h + 1, "ms": hop_ms[h] / K, "frontier_nodes": hop_F[h] / K, "edges": hop_E[h] / K,
                     "GB/s": (alg_bytes[h] / (hop_ms[h] * 1e-3) / 1e9) if hop_ms[h] > 0 else None}
                    for h in range(len(FANOUTS))],
        "all_hops_GBps": alg_bytes.sum() / (hop_ms.sum() * 1e-3) / 1e9,
    }
    traffic_file = os.path.join(ROOT, "profiles", "hop3_dram_bytes_per_launch.json")
    if args.sampler == "uniform" and os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file))["dram_bytes_per_launch"]
        except Exception:
            pass

    # ---- end-to-end with host buffers (e2e) ----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # every transport of SampledBatches.to_host is timed; the line's e2e is the fastest one on this box, the others
        # are kept beside it (which wins depends on the host: PCIe rate against the rate its cores write memory at)
        runs = {}
        # ("mixed" -- 3 compact groups : 1 plain -- measured between compact and hybrid on the 1-GPU box; on request only)
        for tr in (["plain", "compact", "hybrid"] if args.e2e_transport == "all" else [args.e2e_transport]):
            try:
                runs[tr] = run_e2e(thg, plans, streams, host_seeds, B, S, K, W, world, rank, device, cap_n, cap_e, barrier,
                                   transport=tr)
            except Exception as exc:  # e.g. ids beyond int32 for the compact transport
                log(f"[bench] e2e transport {tr} failed: {exc}")
                runs[tr] = None
        good = {k: v for k, v in runs.items() if v}
        if good:
            best = max(good, key=lambda k: good[k]["value"])
            e2e = dict(good[best])
            e2e["other_transports"] = [{kk: v[kk] for kk in ("transport", "value", "ms_per_step", "d2h_bytes_per_step",
                                                              "d2h_GBps_per_gpu", "landing_zone_verified")}
                                       for k, v in good.items() if k != best]

    # ---- the same step with the dedup + insertion-order relabel stage (K7) ------------------------
    with_relabel = None
    if not args.headline_only:
        with_relabel = run_relabel(thg, ptrs, idx, dev_seeds, sampler, B, S, K, W, world, rank, device, barrier)

    clk = clocks.stop()

    # ---- CPU baseline on rank 0, N = 1 only -----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = run_cpu_baseline(ptrs, idx, n, args, oracle_sampler(args.sampler, E, device))

    # ---- BASELINE.json's other single-box configurations, in the same line -----------------------
    walk = hetero = None
    if not args.headline_only:
        res = None
        del plans, plan, dev_seeds, ptrs, idx
        thg.clear_caches()
        torch.cuda.empty_cache()
        walk = measure_walk(args, rank, world, local, device, K=min(K, 5), W=3, with_cpu=False)
        thg.clear_caches()
        torch.cuda.empty_cache()
        hetero = measure_hetero(args, rank, world, local, device, K=K, W=3, with_cpu=False)

    if world > 1:
        dist.barrier()
    if rank == 0:
        extra_launches = sum(x["gpu_launches"] for x in (walk, hetero, with_relabel) if x)
        out = {
            "metric": metric_name(args.sampler), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "config": workload_config(n, E, B, args.scale, args.sampler),
            "seeds_per_sec": seeds_per_s, "edges_per_step_per_gpu": edges / K,
            "roofline": roofline, "serial": serial, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": K * launches_per_step, "gpu_launches_other_measurements": extra_launches,
            "with_relabel": with_relabel,
            "walk_steps_per_sec": walk["value"] if walk else None, "walk": walk,
            "hetero_edges_per_sec": hetero["value"] if hetero else None, "hetero": hetero,
            "clocks": clk, "l2_fetch_granularity": l2_fetch, "to_csc_ms": to_csc_ms, "to_csc_first_call_ms": to_csc_first_ms,
        }
        emit(out)


def run_e2e(thg, plans, streams, host_seeds, B, S, K, W, world, rank, device, cap_n, cap_e, barrier, transport="plain"):
    """Public plan API with HOST buffers.  Per step: pinned seeds -> H2D -> sample (HomogenousSampler.sample_async) ->
    SampledBatches.to_host: the used prefixes of samples / cols / edge_index of every group of 64 batches are packed on
    the device and land in host memory as the reference's i64 vectors (`rows` is arange(S, S + E) for every batch and is
    served from a cached host arange instead of travelling).  Two plans on two streams: step s+1 is sampled while step s
    drains over PCIe.  transport="plain": the packed i64 vectors travel as they are (24 B per edge); "compact": i32 ids
    and one u8 edge count per node travel (9 B per edge) and host threads rebuild the i64 vectors (inside the timed
    region) while the next group is on the bus."""
    from tch_geometric.sharding import reduce_job
    HB = min(B, int(os.environ.get("TCHGEO_BENCH_HB", 64)))  # host landing zone per plan for 64 batches, reused group by group
    # compact / hybrid: a group's staging buffers stay busy until the host threads have rebuilt it, so each plan
    # rotates through three landing zones -- one on the bus, up to two being rebuilt
    ring = 1 if transport == "plain" else int(os.environ.get("TCHGEO_BENCH_RING", 3))
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 2
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    # one group is rebuilt at a time (a single worker per process) with the rank's share of the host cores
    threads = int(os.environ.get("TCHGEO_BENCH_THREADS", max(2, min(32, ncpu // max(local_world, 1)))))
    # "mixed": three of four groups compact, one plain -- the bus carries what the cores cannot write in time
    kinds = ["compact", "compact", "compact", "plain"] if transport == "mixed" else [transport] * ring
    ring = len(kinds)
    try:
        hosts = [[thg.HostBatches(HB, cap_n, cap_e, S, device, fill=0.8, transport=k, threads=threads)
                  for k in kinds] for _ in plans]
    except RuntimeError as e:
        log(f"[bench] could not pin host output buffers: {e}")
        return None
    thg.ops.host_arange(S + cap_e)

    def drain(j):
        """lengths of plan j's pending step, then its packed D2H copies on plan j's stream"""
        res = plans[j].result()
        d2h = res.layer_offsets.nbytes + res.samples_len.nbytes + res.edges_len.nbytes  # already on the host
        with torch.cuda.stream(streams[j]):
            for gi, g0 in enumerate(range(0, B, HB)):
                d2h += res.to_host(hosts[j][gi % ring], g0, min(HB, B - g0))
        return int(res.edges_len.sum()), d2h

    def job(first, count):
        edges, d2h, pending = 0, 0, []
        for i, s in enumerate(range(first, first + count)):
            j = i & 1
            if len(pending) == 2:
                e, d = drain(pending.pop(0))
                edges, d2h = edges + e, d2h + d
            with torch.cuda.stream(streams[j]):
                plans[j].sample_async(host_seeds[s], seed=2000 + s, batch_base=(s * world + rank) * B)  # H2D inside
            pending.append(j)
        while pending:
            e, d = drain(pending.pop(0))
            edges, d2h = edges + e, d2h + d
        for hs in hosts:
            for h in hs:
                h.wait()                     # compact: the last groups' vectors are rebuilt
        return edges, d2h

    job(0, max(min(W, 3), 2))
    barrier()
    cur = torch.cuda.current_stream(device)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    start.record(cur)
    for st in streams:
        st.wait_stream(cur)
    edges, d2h = job(W, K)
    for st in streams:
        cur.wait_stream(st)
    stop.record(cur)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = max(start.elapsed_time(stop), wall_ms)
    # check of the landing zones: every group of the last drained step that is still resident (the last `ring` groups)
    # equals the device result -- first and last batch of the group, all three vectors
    last = plans[(K - 1) & 1]._call
    groups = list(range(0, B, HB))
    ok = True
    for gi in range(max(0, len(groups) - ring), len(groups)):
        hz = hosts[(K - 1) & 1][gi % ring]
        cnt = min(HB, B - groups[gi])
        for i in (0, cnt - 1):
            b = groups[gi] + i
            got = hz.batch(i)
            ne = int(last.edges_len[b, 0])
            ok = ok and bool(torch.equal(got[0], last.samples[0][b, :int(last.samples_len[b, 0])].cpu())
                             and torch.equal(got[2], last.cols[0][b, :ne].cpu())
                             and torch.equal(got[3], last.eidx[0][b, :ne].cpu()))
    ms, edges_all = reduce_job(ms, float(edges), device)
    path = ("HomogenousSampler.sample_async(pinned host seeds) on two plans / two streams + SampledBatches.to_host per "
            "group of 64 batches; rows = arange(S, S+E) is served from a cached host arange (neighbor_sampling.rs:210-218) "
            "and layer offsets / lengths come back with the length table; ")
    if transport == "compact":
        path += (f"transport=compact: tchgeo_pack_transport packs i32 samples / edge positions and one u8 edge count per "
                 f"node, three D2H copies into pinned staging, then tchgeo_host_unpack_transport rebuilds the reference's "
                 f"i64 samples / cols / edge_index in host memory with {threads} threads (inside the timed region)")
    elif transport == "mixed":
        path += (f"transport=mixed: three of every four groups travel compact (i32 ids + u8 edge counts, rebuilt by {threads} "
                 f"host threads inside the timed region), the fourth as packed i64 vectors straight into pinned memory")
    elif transport == "hybrid":
        path += (f"transport=hybrid: edge_index travels as packed i64 straight into its pinned vector; samples travel as "
                 f"i32 and cols as one u8 edge count per node (tchgeo_pack_transport) and tchgeo_host_unpack_transport "
                 f"rebuilds their i64 vectors in host memory with {threads} threads (inside the timed region)")
    else:
        path += "transport=plain: device-side ragged pack, then one D2H per i64 tensor (samples, cols, edge_index) into pinned host buffers"
    del hosts
    return {"value": edges_all / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * S * 8,
            "d2h_bytes_per_step": d2h // K, "ms_per_step": ms / K, "landing_zone_verified": ok,
            "d2h_GBps_per_gpu": d2h / K / (ms / K * 1e-3) / 1e9, "transport": transport,
            "host_bytes_landed_per_step": int(8 * (3 * edges / K + B * S)), "path": path}


def run_relabel(thg, ptrs, idx, dev_seeds, sampler, B, S, K, W, world, rank, device, barrier):
    """The headline step with the K7 stage (frontier dedup + insertion-order relabel of every batch's tree) enqueued
    after the hops.  Per-interval CUDA events (hop launches, then the whole relabel stage) on the launching stream."""
    from tch_geometric.sharding import reduce_job
    plan = thg.HomogenousSampler(ptrs, idx, B, S, FANOUTS, sampler=sampler, relabel=True)
    for s in range(W):
        plan.sample(dev_seeds[s], seed=1000 + s, batch_base=(s * world + rank) * B, timed=True)
    barrier()
    ms = np.zeros(len(FANOUTS) + 1)
    n_samples = n_nodes = edges = 0
    for s in range(W, W + K):
        res = plan.sample(dev_seeds[s], seed=1000 + s, batch_base=(s * world + rank) * B, timed=True)
        ms += res.launch_ms
        n_samples += int(res.samples_len.sum())
        n_nodes += int(res.nodes_len.sum())
        edges += int(res.edges_len.sum())
    step_ms = float(ms.sum())
    step_ms, edges_all = reduce_job(step_ms, float(edges), device)
    peak, peak_src = measured_peak_gbs()
    # algorithmic bytes of the stage: every id is read once (8 B) and gets a local id (8 B); every distinct node is
    # written once (8 B).  The stage's own pair / winner arrays are its overhead, not algorithmic bytes.
    alg = 16.0 * n_samples + 8.0 * n_nodes
    rl_ms = float(ms[-1])
    out = {"value": edges_all / (step_ms * 1e-3), "unit": UNIT, "ms_per_step": step_ms / K,
           "hops_ms_per_step": float(ms[:-1].sum()) / K, "relabel_ms_per_step": rl_ms / K,
           "ids_per_step": n_samples / K, "distinct_nodes_per_step": n_nodes / K,
           "what": "sum of the per-launch CUDA-event intervals of one plan (hop kernels + relabel stage)",
           "roofline": {"bound": "hbm", "kernel": "relabel stage (csrc/relabel.cu, bucketed form, direct-address tables): bk_count / "
                        "bk_tilescan / bk_offsets / bk_scatter_staged / bk_resolve_direct / bk_compact / bk_lookup over all "
                        "batches of the step", "achieved": alg / (rl_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (rl_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_step": alg / K,
                        "byte_model": "16 B per id (8 read + 8 local written) + 8 B per distinct node"},
           "gpu_launches": K * plan.num_launches}
    del plan
    return out


def run_relabel_only(args):
    """--workload relabel: only the with_relabel measurement of the headline configuration (tuning runs)."""
    import tch_geometric as thg
    rank, world, local, device = setup_device()
    B, S, K, W = args.batches, SEEDS_PER_BATCH, args.steps, args.warmup
    ei, n = build_graph(device, args.scale)
    ptrs, idx, _ = thg.to_csc(ei, n)
    del ei
    dev_seeds = torch.stack([torch.from_numpy(synth.seed_batches(n, B, S, first_batch=(s * world + rank) * B)) for s in range(W + K)]).to(device)

    def barrier():
        torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    out = run_relabel(thg, ptrs, idx, dev_seeds, None, B, S, K, W, world, rank, device, barrier)
    out["clocks"] = clocks.stop()
    out["form"] = {k: os.environ[k] for k in ("TCHGEO_RELABEL_DIRECT", "TCHGEO_RELABEL_PERSISTENT", "TCHGEO_RELABEL_WAVES",
                                              "TCHGEO_RELABEL_WAVE_MB") if k in os.environ} or "default (bucketed, direct)"
    if rank == 0:
        emit({"metric": METRIC + "_with_relabel", "n_gpus": world, "steps": K, "warmup": W, "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
              "config": workload_config(n, int(idx.numel()), B, args.scale), **out})


def run_temporal(args):
    """the headline configuration with a TemporalFilter (neighbor_sampling.rs:36-77) on the edges:
    --filter static | relative | dynamic.  Edge timestamps ~ U{0..999}, seed states ~ U{0..999}; windows chosen so that
    about half (static) / a quarter (relative, dynamic) of the edges pass."""
    import tch_geometric as thg
    rank, world, local, device = setup_device()
    B, S, K, W = args.batches, SEEDS_PER_BATCH, args.steps, args.warmup
    ei, n = build_graph(device, args.scale)
    ptrs, idx, _ = thg.to_csc(ei, n)
    del ei
    E = int(idx.numel())
    g = torch.Generator(device=device)
    g.manual_seed(777)
    ts = torch.randint(0, 1000, (E,), generator=g, dtype=torch.int64, device=device)
    mode = {"static": thg.TEMPORAL_SAMPLE_STATIC, "relative": thg.TEMPORAL_SAMPLE_RELATIVE,
            "dynamic": thg.TEMPORAL_SAMPLE_DYNAMIC}[args.filter]
    window = (0, 499) if args.filter == "static" else (0, 249)
    flt = thg.TemporalEdgeFilter(window, ts, True, mode)
    plan = thg.HomogenousSampler(ptrs, idx, B, S, FANOUTS, filter=flt)
    seeds = torch.stack([torch.from_numpy(synth.seed_batches(n, B, S, first_batch=(s * world + rank) * B)) for s in range(W + K)]).to(device)
    states = torch.randint(0, 1000, (W + K, B, S), generator=g, dtype=torch.int64, device=device)
    for s in range(W):
        plan.sample(seeds[s], seed=1000 + s, batch_base=s * B, timed=True, inputs_state=states[s])
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    hop_ms = np.zeros(len(FANOUTS))
    edges = nodes = 0
    deg_sum = 0.0
    for s in range(W, W + K):
        res = plan.sample(seeds[s], seed=1000 + s, batch_base=s * B, timed=True, inputs_state=states[s])
        hop_ms += res.launch_ms
        edges += int(res.edges_len.sum())
        lo = res.layer_offsets
        nodes += int(lo[:, -1, 0].sum())              # frontier nodes of all hops = len(samples) before the last hop
    # untimed: degrees of the last step's frontier nodes (the timestamps the filter has to look at)
    pos = torch.arange(int(plan._call.cap_n[0]), device=device)[None, :]
    f_end = torch.as_tensor(res.layer_offsets[:, -1, 0], device=device)[:, None]
    ids = res.samples[pos < f_end]
    deg_sum = float((ptrs[ids + 1] - ptrs[ids]).sum().item()) * K
    clk = clocks.stop()
    ms = float(hop_ms.sum())
    peak, peak_src = measured_peak_gbs()
    alg = 32.0 * nodes + 48.0 * edges + 8.0 * deg_sum
    cpu = None
    if not args.no_cpu:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        hp, hi, hts = ptrs.cpu().numpy(), idx.cpu().numpy(), ts.cpu().numpy()
        nb = 8
        hs, hst = seeds[W, :nb].cpu().numpy(), states[W, :nb].cpu().numpy()
        t0 = time.perf_counter()
        tot = 0
        for b in range(nb):
            o = O.neighbor_sampling_homogenous(hp, hi, hs[b], FANOUTS, filter=dict(mode=mode, forward=True, window=window, timestamps=hts, inputs_state=hst[b]),
                                               rng_mode=O.RNG_XOSHIRO, seed=b)
            tot += o[1].size
        dt = time.perf_counter() - t0
        cpu = {"value": tot / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{nb} batches x {S} seeds, one thread like the reference ({dt:.1f} s)"}
    emit({"metric": f"{METRIC}_temporal_{args.filter}", "value": edges / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": K,
          "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "int64", "data": "synthetic",
          "config": {"workload": workload_config(n, E, B, args.scale)["workload"] + f", TemporalFilter {args.filter} "
                     f"window {window}, forward, edge timestamps and seed states ~ U{{0..999}}",
                     "l2_policy": "inputs larger than L2; seeds and states differ every step"},
          "edges_per_step": edges / K, "frontier_nodes_per_step": nodes / K, "frontier_degree_sum_per_step": deg_sum / K,
          "roofline": {"bound": "hbm", "kernel": "hop_filtered_kernel<UNIFORM>, all hops of the step",
                       "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                       "traffic": None, "peak_source": peak_src, "per_hop_ms": [float(x) / K for x in hop_ms],
                       "byte_model": "32 B per frontier node (id, colptr pair, state) + 48 B per sampled edge (gather, four "
                                     "outputs, state) + 8 B per timestamp of the frontier's neighbourhoods"},
          "cpu_baseline": cpu, "e2e": None, "gpu_launches": K * plan.num_launches, "clocks": clk})


def run_cpu_baseline(ptrs, idx, n, args, osampler=None):
    cores = os.cpu_count() or 1
    hp, hi = ptrs.cpu().numpy(), idx.cpu().numpy()
    r1, e1, t1 = cpu_sampling_rate(hp, hi, n, 8, 1, first_batch=10_000_000, sampler=osampler)
    nb = int(max(cores * 4, min(4096, 5.0 * cores / (t1 / 8))))  # ~5 s of wall time on all cores
    rT, eT, tT = cpu_sampling_rate(hp, hi, n, nb, cores, first_batch=10_000_100, sampler=osampler)
    log(f"[bench] cpu baseline: 1 thread {r1 / 1e6:.2f} M edges/s, {cores} threads {rT / 1e6:.2f} M edges/s")
    return {"value": rT, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{nb} batches x {SEEDS_PER_BATCH} seeds on {cores} threads ({tT:.1f} s); "
                      f"1 thread: {r1:.0f} edges/s on 8 batches ({t1:.1f} s)",
            "value_1_thread": r1}


# ---------------------------------------------------------------------------------------------
# secondary workloads (BASELINE.json configs[2] and configs[3]); same JSON shape, own metric names
# ---------------------------------------------------------------------------------------------
def run_walk(args):
    rank, world, local, device = setup_device()
    out = measure_walk(args, rank, world, local, device, args.steps, args.warmup, with_cpu=not args.no_cpu)
    if rank == 0:
        emit(out)


def measure_walk(args, rank, world, local, device, K, W, with_cpu):
    """configs[2]: node2vec random_walk, walk_length 80, p=1, q=0.5, 10 walks per node, walkers sharded (strong scaling)."""
    import tch_geometric as thg
    import torch.distributed as dist
    from tch_geometric.sharding import reduce_job, shard_range
    ei, n = build_graph(device, args.scale)
    rp, ci, _ = thg.to_csr(ei, n)
    del ei
    torch.cuda.empty_cache()
    L, P, Q = 80, 1.0, 0.5
    total_walkers = n * 10 if not args.walkers else int(args.walkers)  # --walkers: profiling runs only
    w0, w1 = shard_range(total_walkers, rank, world)          # strong scaling: the job is fixed
    start = (torch.arange(w0, w1, device=device, dtype=torch.int64) // 10)  # arange(N).repeat_interleave(10)
    walks = None
    for s in range(max(W, 2)):
        # the result is held across iterations exactly as in the timed loop, so that BOTH output buffers the loop
        # alternates between exist before it starts (a cudaMalloc of the second one -- 15.9 GB / world -- used to land
        # inside the timed region: 5-25 % of a 5-step measurement, and the source of its run-to-run spread)
        walks = thg.random_walk(rp, ci, start, L, P, Q, seed=s, walker_base=w0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps_taken, attempts = 0, 0
    e0.record()
    for s in range(K):
        walks, att = thg.random_walk(rp, ci, start, L, P, Q, seed=100 + s, walker_base=w0, return_attempts=True)
        attempts += att
        if s == K - 1:
            e1.record()
            steps_taken = int((walks[:, 1:] >= 0).sum().item())
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    ms, steps_all = reduce_job(ms, float(steps_taken) * K, device)
    _, att_all = reduce_job(0.0, float(attempts), device)
    value = steps_all / (ms * 1e-3)
    peak, peak_src = measured_peak_gbs()
    deg_mean = ci.numel() / n
    bytes_per_attempt = 24 + 8 * int(np.ceil(np.log2(max(deg_mean, 2))))
    alg = (att_all * bytes_per_attempt + steps_all * 8) / world  # per-GPU bytes over the timed region
    achieved = alg / (ms * 1e-3) / 1e9
    cpu = None
    if rank == 0 and world == 1 and with_cpu:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        sub = 1_000_000 if args.scale == 1.0 else min(100_000, total_walkers)
        hrp, hci = rp.cpu().numpy(), ci.cpu().numpy()
        st = start[:sub].cpu().numpy()
        t0 = time.perf_counter()
        wk, _ = O.random_walk_mt(hrp, hci, st, L, P, Q, rng_mode=O.RNG_XOSHIRO, seed=1, num_threads=cores)
        dt = time.perf_counter() - t0
        cpu = {"value": float((wk[:, 1:] >= 0).sum()) / dt, "unit": "steps/s", "cores": cores, "kind": "port",
               "sample": f"{sub} walkers x {L} steps on {cores} threads ({dt:.1f} s)"}
    del walks
    return ({"metric": "node2vec_walk_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": K,
              "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "int64", "data": "synthetic",
              "config": {"workload": f"products-shaped synthetic graph (N={n}, E={ci.numel()}), random_walk "
                                     f"walk_length={L}, p={P}, q={Q}, {total_walkers} walkers (10 per node)",
                         "l2_policy": "inputs larger than L2 (col_indices 495 MB, walks 15.9 GB)",
                         "parallelism": "walkers sharded over ranks, CSR replicated, no collective"},
              "attempts_per_step": att_all / max(steps_all, 1.0),
              "roofline": {"bound": "hbm", "kernel": "walk_kernel",
                           "achieved": achieved, "peak": peak, "unit": "GB/s",
                           "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                           "algorithmic_bytes_per_attempt": bytes_per_attempt,
                           "note": "algorithmic bytes follow SURVEY 8(d) (reference algorithm: neighbour gather + row_ptrs "
                                   "pair + binary search per attempt); the kernel skips the searches the uniform already decides "
                                   "(about half of them at p=1, q=0.5)"},
              "cpu_baseline": cpu, "e2e": None, "gpu_launches": K, "clocks": clk})


def run_negative(args):
    """negative_sample_neighbors_homogenous on the products-shaped CSR, 1 Mi inputs x 5 negatives."""
    import tch_geometric as thg
    rank, world, local = dist_env()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    ei, n = build_graph(device, args.scale)
    rp, ci, _ = thg.to_csr(ei, n)
    del ei
    torch.cuda.empty_cache()
    S, NEG, TRY = (1 << 20 if args.scale == 1.0 else 1 << 14), 5, 5
    K, W = args.steps, args.warmup
    g = torch.Generator(device=device)
    g.manual_seed(99)
    inputs = [torch.randint(0, n, (S,), generator=g, dtype=torch.int64, device=device) for _ in range(W + K)]
    out = None
    for s in range(max(W, 2)):   # (result held as in the timed loop: both sets of output buffers exist before it)
        out = thg.negative_sample_neighbors_homogenous(rp, ci, (n, n), inputs[s % len(inputs)], NEG, TRY, seed=s)
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    edges = 0
    e0.record()
    for s in range(W, W + K):
        out = thg.negative_sample_neighbors_homogenous(rp, ci, (n, n), inputs[s], NEG, TRY, seed=100 + s)
        edges += out[1].numel()
    e1.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    value = edges / (ms * 1e-3)
    cpu = None
    if not args.no_cpu:
        from oracle import oracle as O
        hrp, hci = rp.cpu().numpy(), ci.cpu().numpy()
        sub = inputs[W][: min(S, 200_000)].cpu().numpy()
        t0 = time.perf_counter()
        o = O.negative_sample_neighbors_homogenous(hrp, hci, (n, n), sub, NEG, TRY, rng_mode=O.RNG_XOSHIRO, seed=1)
        dt = time.perf_counter() - t0
        cpu = {"value": o[1].size / dt, "unit": "negatives/s", "cores": 1, "kind": "port",
               "sample": f"{sub.size} inputs x {NEG} negatives, single thread like the reference ({dt:.1f} s)"}
    deg_mean = ci.numel() / n
    alg = edges * (16 + 8 * int(np.ceil(np.log2(max(deg_mean, 2)))) + 24)  # row_ptrs pair + search + (rows, cols, sample)
    peak, peak_src = measured_peak_gbs()
    emit({"metric": "negative_samples_per_sec", "value": value, "unit": "negatives/s", "n_gpus": 1, "steps": K, "warmup": W,
          "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
          "data": "synthetic",
          "config": {"workload": f"products-shaped synthetic graph (N={n}, E={ci.numel()}), negative_sample_neighbors_homogenous, "
                                 f"{S} inputs, num_neg={NEG}, try_count={TRY}",
                     "l2_policy": "inputs differ every step; col_indices is 495 MB"},
          "roofline": {"bound": "hbm", "kernel": "whole call (draw + 2 compactions + relabel)", "achieved": alg / (ms * 1e-3) / 1e9,
                       "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                       "peak_source": peak_src},
          "cpu_baseline": cpu, "e2e": None, "gpu_launches": K * 12, "clocks": clk})


def run_tempo_walk(args):
    """tempo_random_walk (random_walk.rs:80-158) on the products-shaped graph: one walker per node,
    walk_length 20, edge timestamps ~ U{0..999}, start timestamps ~ U{0..999}, window (0, 500).  The window is relative
    to the walker's START timestamp (:101-103), so about half of a node's edges pass at every step (less for late
    starts) until a walker reaches a node without a passing edge."""
    import tch_geometric as thg
    rank, world, local, device = setup_device()
    ei, n = build_graph(device, args.scale)
    rp, ci, _ = thg.to_csr(ei, n)
    del ei
    torch.cuda.empty_cache()
    E, L, K, W = int(ci.numel()), 20, args.steps, args.warmup
    g = torch.Generator(device=device)
    g.manual_seed(4242)
    ets = torch.randint(0, 1000, (E,), generator=g, dtype=torch.int64, device=device)
    nts = torch.randint(0, 1000, (n,), generator=g, dtype=torch.int64, device=device)
    start = torch.arange(n, dtype=torch.int64, device=device)
    sts = torch.randint(0, 1000, (n,), generator=g, dtype=torch.int64, device=device)
    window = (0, 500)
    walks = wts = None
    for s in range(max(W, 2)):   # (results held across iterations, as in the timed loop: both output sets exist before it)
        walks, wts = thg.tempo_random_walk(rp, ci, nts, ets, start, sts, L, window, seed=s)
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(K):
        walks, wts = thg.tempo_random_walk(rp, ci, nts, ets, start, sts, L, window, seed=100 + s)
    e1.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1) / K
    # work of the last call: every position that is followed by an attempt (a live walker looks at ALL neighbours of
    # its current node: col_indices + edge_timestamps, 16 B each, plus the row_ptrs pair), every output cell written
    deg = rp[1:] - rp[:-1]
    live = walks[:, :-1] >= 0
    steps_taken = int((walks[:, 1:] >= 0).sum())
    looked = int(deg[walks[:, :-1][live]].sum())
    attempts = int(live.sum())
    alg = 16.0 * looked + 16.0 * attempts + 16.0 * walks.numel()
    peak, peak_src = measured_peak_gbs()
    cpu = None
    if not args.no_cpu:
        from oracle import oracle as O
        sub = min(n, 200_000)
        hrp, hci, hn, he = rp.cpu().numpy(), ci.cpu().numpy(), nts.cpu().numpy(), ets.cpu().numpy()
        t0 = time.perf_counter()
        wk, _ = O.tempo_random_walk(hrp, hci, hn, he, start[:sub].cpu().numpy(), sts[:sub].cpu().numpy(), L, window,
                                    rng_mode=O.RNG_XOSHIRO, seed=1)
        dt = time.perf_counter() - t0
        cpu = {"value": float((wk[:, 1:] >= 0).sum()) / dt, "unit": "steps/s", "cores": 1, "kind": "port",
               "sample": f"{sub} walkers x {L} on 1 thread ({dt:.1f} s)"}
    emit({"metric": "tempo_walk_steps_per_sec", "value": steps_taken / (ms * 1e-3), "unit": "steps/s", "n_gpus": world,
          "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "int64", "data": "synthetic",
          "config": {"workload": f"products-shaped synthetic graph (N={n}, E={E}), tempo_random_walk walk_length={L}, window "
                                 f"{window} relative to the start timestamp, {n} walkers (one per node), timestamps U{{0..999}}",
                     "l2_policy": "inputs larger than L2 (col_indices + edge_timestamps 990 MB)"},
          "steps_taken_per_call": steps_taken, "neighbours_examined_per_call": looked,
          "roofline": {"bound": "hbm", "kernel": "tempo_walk_kernel", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak,
                       "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                       "byte_model": "16 B per neighbour examined (id + edge timestamp) + 16 B row_ptrs pair per attempt + "
                                     "16 B per output cell (walks, walk_timestamps)"},
          "cpu_baseline": cpu, "e2e": None, "gpu_launches": K, "clocks": clk})


def run_gather(args):
    """x[samples] for the samples of one sampling step (products-shaped graph, 100 f32 features)."""
    import tch_geometric as thg
    rank, world, local = dist_env()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    ei, n = build_graph(device, args.scale)
    ptrs, idx, _ = thg.to_csc(ei, n)
    del ei
    torch.cuda.empty_cache()
    D, B, S, K, W = 100, min(args.batches, 32), SEEDS_PER_BATCH, args.steps, args.warmup
    x = torch.randn(n, D, device=device)
    plan = thg.HomogenousSampler(ptrs, idx, B, S, FANOUTS)
    index = []
    for s in range(W + K):  # every step gathers the samples of a different sampling step (inputs differ every step)
        res = plan.sample(torch.from_numpy(synth.seed_batches(n, B, S, first_batch=s * B)).to(device), seed=s)
        index.append(torch.cat([res.samples[b, :int(res.samples_len[b])] for b in range(B)]))
    cap_rows = max(int(i.numel()) for i in index)
    flat = torch.empty(cap_rows * D, device=device)               # one preallocated output, reused every step
    view = lambda s: flat[: index[s].numel() * D].view(-1, D)
    for s in range(W):
        thg.gather_rows(x, index[s], out=view(s))
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rows = 0
    e0.record()
    for s in range(W, W + K):
        out = thg.gather_rows(x, index[s], out=view(s))           # validating call: one error read-back per call
        rows += out.shape[0]
    e1.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for s in range(W, W + K):                                     # kernel only: asynchronous calls back to back
        thg.gather_rows(x, index[s], out=view(s), validate=False)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / K
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in range(W, W + K):
        ref = x[index[s]]
    t1.record()
    torch.cuda.synchronize()
    assert torch.equal(ref, out)
    alg = rows * (8 + 2 * D * 4)
    alg_launch = alg / K
    peak, peak_src = measured_peak_gbs()
    cpu = None
    if not args.no_cpu:
        hx, hi = x.cpu().numpy(), index[W][:2_000_000].cpu().numpy()
        t = time.perf_counter()
        _ = hx[hi]
        dt = time.perf_counter() - t
        cpu = {"value": hi.size / dt, "unit": "rows/s", "cores": 1, "kind": "port",
               "sample": f"numpy fancy indexing of {hi.size} rows x {D} f32 ({dt:.2f} s)"}
    emit({"metric": "gathered_feature_rows_per_sec", "value": rows / (ms * 1e-3), "unit": "rows/s", "n_gpus": 1, "steps": K,
          "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f32 rows (byte copy)", "data": "synthetic",
          "config": {"workload": f"x[samples]: x = [{n}, {D}] f32, index = samples of {B} batches of the 3-hop {FANOUTS} "
                                 f"sampling step ({rows // K} rows per step)",
                     "l2_policy": "x is 980 MB, output 2 x larger than L2; a different index every step"},
          "roofline": {"bound": "hbm", "kernel": "gather_rows_kernel<uint4>", "achieved": alg_launch / (kernel_ms * 1e-3) / 1e9,
                       "peak": peak, "unit": "GB/s", "frac": alg_launch / (kernel_ms * 1e-3) / 1e9 / peak,
                       "launch_ms": kernel_ms, "traffic": 16928487000, "traffic_source": "profiles/r1_gather_ncu_metrics.csv",
                       "peak_source": peak_src, "algorithmic_bytes_per_row": 8 + 2 * D * 4},
          "torch_index_ms_per_step": t0.elapsed_time(t1) / K,
          "cpu_baseline": cpu, "e2e": None, "gpu_launches": K, "clocks": clk})


def run_hetero(args):
    rank, world, local, device = setup_device()
    out = measure_hetero(args, rank, world, local, device, args.steps, args.warmup, with_cpu=not args.no_cpu)
    if rank == 0:
        emit(out)


def measure_hetero(args, rank, world, local, device, K, W, with_cpu):
    """configs[3]: ogbn-mag-shaped heterogeneous sampling, fanouts [10,10] per relation, 1024 paper seeds per batch,
    seed batches sharded over the ranks (weak scaling), all four CSCs replicated."""
    import tch_geometric as thg
    import torch.distributed as dist
    from tch_geometric.sharding import reduce_job
    counts, edges = synth.mag_like(device, scale=args.scale)
    node_types = list(counts)
    edge_types = list(edges)
    cp, ri = {}, {}
    for et, ei in edges.items():
        p_, i_, _ = thg.to_csc(ei, (counts[et[0]], counts[et[2]]))
        cp[thg.rel_key(et)], ri[thg.rel_key(et)] = p_, i_
    B, S, H = args.batches, SEEDS_PER_BATCH, 2
    nn = {thg.rel_key(et): [10, 10] for et in edge_types}
    plan = thg.HeterogenousSampler(node_types, edge_types, cp, ri, B, {"paper": S}, nn, H)

    def seeds(step):
        out = np.empty((B, S), dtype=np.int64)
        for b in range(B):
            out[b] = np.random.default_rng(1234 + (step * world + rank) * B + b).choice(counts["paper"], S, replace=False)
        return torch.from_numpy(out).to(device)

    all_seeds = [seeds(s) for s in range(W + K)]
    for s in range(W):
        plan.sample({"paper": all_seeds[s]}, seed=1000 + s, batch_base=(s * world + rank) * B)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    edges_n = 0
    launch_ms = None
    alg_bytes = 0.0
    src_of = [node_types.index(et[0]) for et in edge_types]
    dst_of = [node_types.index(et[2]) for et in edge_types]
    seeds_of = np.array([S if t == "paper" else 0 for t in node_types], dtype=np.float64)
    e0.record()
    for s in range(W, W + K):
        plan.sample({"paper": all_seeds[s]}, seed=1000 + s, batch_base=(s * world + rank) * B, timed=True)
        edges_n += int(plan.edges_len.sum())
        launch_ms = plan.launch_ms if launch_ms is None else launch_ms + plan.launch_ms
        # algorithmic bytes of the step, SURVEY 8(d): 24 B per frontier node + 40 B per sampled edge of every launch.
        # E[r,h] from the layer offsets; frontier of hop 0 = the seeds of the dst type, of hop 1 = the nodes the dst
        # type gained during hop 0 (neighbor_sampling.rs:345-348).
        lo, el = plan.layer_offsets, plan.edges_len  # [B, R, H, 3], [B, R]
        e_h0 = (lo[:, :, 1, 1] - lo[:, :, 0, 1]).sum(axis=0).astype(np.float64)  # [R]
        e_h1 = (el - lo[:, :, 1, 1]).sum(axis=0).astype(np.float64)
        gained = np.zeros(len(node_types))
        for r, st in enumerate(src_of):
            gained[st] += e_h0[r]
        f_h0 = np.array([seeds_of[d] * B for d in dst_of])
        f_h1 = np.array([gained[d] for d in dst_of])
        alg_bytes += 24.0 * (f_h0.sum() + f_h1.sum()) + 40.0 * (e_h0.sum() + e_h1.sum())
    e1.record()
    torch.cuda.synchronize()
    serial_ms = e0.elapsed_time(e1)
    # value: the same K steps with two plans on two streams (sample_async / result): the host enqueues step s+1 before it
    # waits for the lengths of step s
    plans = [plan, thg.HeterogenousSampler(node_types, edge_types, cp, ri, B, {"paper": S}, nn, H)]
    streams = [torch.cuda.Stream(device), torch.cuda.Stream(device)]

    def pipelined(first, count):
        total, pending = 0, []
        for i, s in enumerate(range(first, first + count)):
            j = i & 1
            if len(pending) == 2:
                total += int(plans[pending.pop(0)].result().edges_len.sum())
            with torch.cuda.stream(streams[j]):
                plans[j].sample_async({"paper": all_seeds[s]}, seed=1000 + s, batch_base=(s * world + rank) * B)
            pending.append(j)
        while pending:
            total += int(plans[pending.pop(0)].result().edges_len.sum())
        return total

    pipelined(0, min(W, len(all_seeds)))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    cur = torch.cuda.current_stream(device)
    e0.record(cur)
    for st in streams:
        st.wait_stream(cur)
    edges_p = pipelined(W, K)
    for st in streams:
        cur.wait_stream(st)
    e1.record(cur)
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    assert edges_p == edges_n
    ms, edges_all = reduce_job(ms, float(edges_n), device)
    cpu = None
    if rank == 0 and world == 1 and with_cpu:
        from concurrent.futures import ThreadPoolExecutor
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        hcp = {k: v.cpu().numpy() for k, v in cp.items()}
        hri = {k: v.cpu().numpy() for k, v in ri.items()}
        hs = all_seeds[0].cpu().numpy()
        nb = min(B, cores * 16)

        def one(b):
            out = O.neighbor_sampling_heterogenous(node_types, edge_types, hcp, hri, {"paper": hs[b]}, nn, H,
                                                   rng_mode=O.RNG_XOSHIRO, seed=b)
            return sum(len(v) for v in out[1].values())
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            tot = sum(ex.map(one, range(nb)))
        dt = time.perf_counter() - t0
        cpu = {"value": tot / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{nb} batches x {S} paper seeds on {cores} threads ({dt:.1f} s)"}
    launches = plan.num_launches
    del plan, plans
    return ({"metric": "sampled_edges_per_sec_hetero_mag_10_10", "value": edges_all / (ms * 1e-3), "unit": UNIT,
              "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
              "config": {"workload": f"ogbn-mag-shaped synthetic hetero graph {counts}, 4 relations, "
                                     f"neighbor_sampling_heterogenous fanouts [10,10] per relation, {S} paper seeds/batch, "
                                     f"{B} batches/step", "parallelism": "seed batches sharded, CSCs replicated"},
              "launch_ms": [float(x) / K for x in launch_ms], "edges_per_step_per_gpu": edges_n / K,
              "serial_ms_per_step": serial_ms / K,
              "roofline": {"bound": "hbm", "kernel": "hop_kernel<UNIFORM>, all launches of the step (one per hop and relation "
                           "with a non-empty frontier)", "achieved": alg_bytes / (float(launch_ms.sum()) * 1e-3) / 1e9,
                           "peak": measured_peak_gbs()[0], "unit": "GB/s",
                           "frac": alg_bytes / (float(launch_ms.sum()) * 1e-3) / 1e9 / measured_peak_gbs()[0],
                           "traffic": None, "algorithmic_bytes_per_step": alg_bytes / K,
                           "kernel_ms_per_step": float(launch_ms.sum()) / K},
              "cpu_baseline": cpu, "e2e": None, "gpu_launches": K * launches, "clocks": clk})


def run_partitioned(args):
    """configs[4]: papers100M-shaped graph, CSC range-partitioned over the ranks, 3-hop [15,10,5] sampling with the
    frontier exchanged per hop (requests out, sampled neighbours back).  --protocol fixed (default): both exchanges are
    NVLink peer-memory stores into fixed per-pair segments, counts stay on the device, no host synchronisation inside a
    step; --protocol legacy: round 1's count matrix + host read per hop (NCCL all-to-alls or peer stores)."""
    import tch_geometric as thg
    import torch.distributed as dist
    from tch_geometric.partitioned import DistComm, PartitionedPlan, PartitionedPlanF, PartitionedPlanGroups, SingleComm
    from tch_geometric.sharding import reduce_job
    rank, world, local, device = setup_device()
    part, n, e_total, cols_rank = synth.papers_partition(thg, rank, world, device, args.scale)
    B, S, K, W = args.batches, SEEDS_PER_BATCH, args.steps, args.warmup
    comm = DistComm() if world > 1 else SingleComm()
    n_groups = args.groups if args.groups > 0 else 1   # measured on 8 B200s: 1 group 3.50 ms, 2 groups 4.23, 4 groups 5.13
    if args.protocol == "fixed" and n_groups > 1:
        ps = PartitionedPlanGroups(part, B, S, FANOUTS, comm=comm if world > 1 else None, groups=n_groups, slack=args.slack)
    elif args.protocol == "fixed":
        ps = PartitionedPlanF(part, B, S, FANOUTS, comm=comm if world > 1 else None, world=world, rank=rank, slack=args.slack)
    else:
        ps = PartitionedPlan(part, B, S, FANOUTS, comm=comm, groups=args.groups if args.groups > 0 else None)
    host_seeds = torch.empty((W + K, B, S), dtype=torch.int64).pin_memory()
    for s in range(W + K):
        host_seeds[s] = torch.from_numpy(synth.seed_batches(n, B, S, first_batch=(s * world + rank) * B))
    seeds = host_seeds.to(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(W):
        ps.sample(seeds[s], seed=1000 + s, batch_base=(s * world + rank) * B)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    edges_n = 0
    F_hop = np.zeros(len(FANOUTS))          # frontier nodes per hop, summed over the timed steps
    e0.record()
    for s in range(W, W + K):
        out = ps.sample(seeds[s], seed=1000 + s, batch_base=(s * world + rank) * B)
        edges_n += int(out.edges_len.sum())
        nl = out._node_len.sum(axis=1).astype(np.float64)      # [H+1] len(samples) after h hops
        F_hop += np.diff(np.concatenate([[0.0], nl[:-1]]))
    e1.record()
    torch.cuda.synchronize()
    ps.profile = {}   # untimed extra pass: device time per phase (CUDA events between the phases)
    for s in range(W, W + K):
        ps.sample(seeds[s], seed=1000 + s, batch_base=(s * world + rank) * B)
    phases = {k: v / ps.profile["calls"] for k, v in ps.profile.items() if k != "calls"}
    ps.profile = None
    ms, edges_all = reduce_job(e0.elapsed_time(e1), float(edges_n), device)

    # ---- e2e: pinned host seeds -> H2D -> sample -> packed D2H of samples / cols / edge_index ----------------------
    e2e = None
    if not args.no_e2e:
        HB = min(B, 64)
        host = thg.HostBatches(HB, ps.cap_n, ps.cap_e, S, device, fill=0.8)
        thg.ops.host_arange(S + ps.cap_e)

        def e2e_step(s):
            out = ps.sample(host_seeds[s], seed=2000 + s, batch_base=(s * world + rank) * B)
            d = 0
            for g0 in range(0, B, HB):
                d += out.to_host(host, g0, min(HB, B - g0))
            return int(out.edges_len.sum()), d
        e2e_step(0)
        barrier()
        t0 = time.perf_counter()
        ee, dd = 0, 0
        for s in range(W, W + K):
            a, d = e2e_step(s)
            ee, dd = ee + a, dd + d
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        e2e_ms, ee_all = reduce_job(e2e_ms, float(ee), device)
        e2e = {"value": ee_all / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * S * 8, "d2h_bytes_per_step": dd // K,
               "ms_per_step": e2e_ms / K,
               "path": "plan.sample(pinned host seeds) + PartitionedBatches.to_host (device-side ragged pack, one D2H per tensor "
                       "and group of 64 batches; rows = arange served from the host)"}
    clk = clocks.stop()

    # ---- CPU baseline (N = 1 only): the oracle port on this rank's 13.9 M-column share ----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = run_cpu_baseline(part.ptrs, part.indices, n, args, None)

    peak, peak_src = measured_peak_gbs()
    nvlink_gbs = 600.0   # all_to_all_single with static buffers on this box class (profiles/r1_a2a_probe_8gpu.txt)
    F_tot = float(F_hop.sum())
    req_b = 16.0 * F_tot / K                                   # per rank per step: 16-byte request rows
    ans_b = float(sum(8.0 * k * f for k, f in zip(FANOUTS, F_hop))) / K   # 2k int32 words per frontier node of hop h
    remote = (world - 1) / world
    exch_ms = sum(v for k_, v in phases.items() if "barrier" in k_ or "a2a" in k_)
    kern_ms = sum(v for k_, v in phases.items() if not ("barrier" in k_ or "a2a" in k_))
    alg = 24.0 * F_tot / K + 40.0 * edges_n / K                # what the replicated kernel moves for the same step
    if rank == 0:
        emit({"metric": "sampled_edges_per_sec_3hop_15_10_5_partitioned_csc", "value": edges_all / (ms * 1e-3),
              "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
              "config": {"workload": f"papers100M-shaped synthetic graph (N={n}, E={e_total}; {cols_rank} columns per rank), "
                                     f"CSC range-partitioned over {world} rank(s), fanouts {FANOUTS}, {S} seeds/batch, "
                                     f"{B} batches/step/rank",
                         "protocol": args.protocol, "pipelined_batch_groups": getattr(ps, "num_groups", 1),
                         "parallelism": "column-range partition; per hop the frontier goes to the owners and the sampled "
                                        "neighbours come back"
                                        + (" as NVLink peer-memory stores into fixed per-pair segments (no host sync, no "
                                           "count matrix)" if args.protocol == "fixed" else " (count matrix + host read per hop)"),
                         "l2_policy": "inputs larger than L2 (row_indices 1.6 GB per rank; GBs of outputs per step); seeds "
                                      "differ every step"},
              "phase_ms_per_step_rank0": phases, "kernels_ms_per_step_rank0": kern_ms, "exchange_wait_ms_per_step_rank0": exch_ms,
              "exchange_bytes_per_step_per_rank": {"requests": req_b, "answers": ans_b, "leaving_the_gpu": (req_b + ans_b) * remote},
              "roofline": {"bound": "hbm", "kernel": "whole step (scatter + put, serve, finish per hop)",
                           "achieved": alg / (ms / K * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": alg / (ms / K * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                           "algorithmic_bytes_per_step": alg,
                           "byte_model": "24 B per frontier node + 40 B per sampled edge (SURVEY 8d: the bytes of the replicated "
                                         "step; request and answer rows are protocol overhead on top)",
                           "exchange": {"bytes_leaving_per_step": (req_b + ans_b) * remote, "nvlink_GBps_reference": nvlink_gbs,
                                        "ms_at_reference_rate": (req_b + ans_b) * remote / nvlink_gbs / 1e6,
                                        "note": "the stores ride inside the scatter/put and serve kernels; the barrier "
                                                "phases are what is left of the exchange as waiting time"}},
              "cpu_baseline": cpu, "e2e": e2e,
              "gpu_launches": K * len(FANOUTS) * (4 if args.protocol == "fixed" else 8), "clocks": clk})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sampling", choices=["sampling", "walk", "hetero", "partitioned", "negative", "gather", "relabel", "temporal", "tempo_walk"],
                    help="sampling = headline (configs[1]); walk = configs[2]; hetero = configs[3]")
    ap.add_argument("--batches", type=int, default=256, help="seed batches per step per GPU")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (debugging only)")
    ap.add_argument("--ref-batches", type=int, default=0, help="batches per step of the reference arm (0 = --batches)")
    ap.add_argument("--headline-only", action="store_true",
                    help="sampling workload: skip the walk / hetero / relabel measurements attached to the default line")
    ap.add_argument("--l2-fetch", type=int, default=0, help="set cudaLimitMaxL2FetchGranularity (0 = leave)")
    ap.add_argument("--sampler", default="uniform", choices=["uniform", "replace", "weighted"],
                    help="sampling workload: neighbour sampler (uniform = the headline configuration)")
    ap.add_argument("--filter", default="static", choices=["static", "relative", "dynamic"],
                    help="temporal workload: TemporalFilter mode")
    ap.add_argument("--groups", type=int, default=0,
                    help="partitioned workload: batch groups pipelined on separate streams (0 = default: 1)")
    ap.add_argument("--protocol", default="fixed", choices=["fixed", "legacy"],
                    help="partitioned workload: exchange protocol (fixed = device-only fixed segments; legacy = round 1)")
    ap.add_argument("--slack", type=float, default=1.5,
                    help="partitioned workload, fixed protocol: segment size as a multiple of the mean per-pair load")
    ap.add_argument("--walkers", type=int, default=0, help="walk workload: number of walkers (0 = 10 per node)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-transport", choices=["all", "plain", "compact", "hybrid", "mixed"], default="all",
                    help="SampledBatches.to_host transport(s) to time for the e2e number")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "walk":
        run_walk(args)
    elif args.workload == "hetero":
        run_hetero(args)
    elif args.workload == "partitioned":
        run_partitioned(args)
    elif args.workload == "negative":
        run_negative(args)
    elif args.workload == "gather":
        run_gather(args)
    elif args.workload == "relabel":
        run_relabel_only(args)
    elif args.workload == "temporal":
        run_temporal(args)
    elif args.workload == "tempo_walk":
        run_tempo_walk(args)
    else:
        run_ours(args)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
