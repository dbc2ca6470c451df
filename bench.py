#!/usr/bin/env python
"""bench.py -- collision-checked edges/s and roadmap build time of the PRM learning phase.

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): 4 Franka
Pandas on a circle of radius 0.665 m facing the centre above a ground plane
(experiments/pybullet/exp_scalability.py:197-213), 1 M collision-free joint-space samples
per arm, k = 10 nearest candidates under the reference's FK metric
(robots/panda.py:248-254), every candidate edge checked at 50 interpolation steps
(local_step 0.02).  One step = the whole learning phase for all arms:
sample -> config collision -> compaction -> FK features -> k-NN scan + fp64 re-rank
-> edge collision, all device resident.  With N GPUs the total work is fixed (strong
scaling): every rank draws a block of the candidate stream, answers one spatial slice of the
neighbour search and checks one row block of the candidate edges; the pieces are exchanged
with NCCL all-gather.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "collision_checked_edges_per_s"
UNIT = "edges/s"


def scene(n_arms, radius=0.665):
    """Base poses of exp_scalability.calculate_base_positions (reference :197-210)."""
    from mmmp_b200 import core
    bases = []
    for i in range(n_arms):
        ang = i * 2 * np.pi / n_arms
        yaw = ang + np.pi
        bases.append(core.base_matrix((radius * np.cos(ang), radius * np.sin(ang), 0.01),
                                      (0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2))))
    return bases, [core.plane((0, 0, 1), 0.0)]


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------- CPU baseline
def _cpu_sample(seed, n_nodes, k, n_steps, step):
    """The reference's learning phase restated on the CPU for ONE arm on a bounded sample:
    rejection-sample n_nodes free configs (hull checker = restatement of the pybullet
    predicates), heap-loop semantics k-NN under the FK metric (vectorised float64), then
    is_collision_free_edge on every candidate.  Returns (edges_checked, seconds)."""
    from oracle import chain
    from oracle.hull import PandaHullChecker
    from mmmp_b200 import core  # only for base_matrix/plane helpers (host math, no GPU)
    bases, obstacles = scene(4)
    ck = PandaHullChecker([bases[0]], obstacles)
    model_limits = np.array([(-2.8973, 2.8973), (-1.7628, 1.7628), (-2.8973, 2.8973), (-3.0718, -0.0698),
                             (-2.8973, 2.8973), (-0.0175, 3.7525), (-2.8973, 2.8973)])
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    nodes = np.zeros((0, 7))
    while len(nodes) < n_nodes:
        q = np.round(rng.uniform(model_limits[:, 0], model_limits[:, 1], (n_nodes, 7)), 2)
        nodes = np.concatenate([nodes, q[~ck.config_collision(q, 0)]])
    nodes = nodes[:n_nodes]
    P = chain.panda_positions(nodes, bases[0])
    D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1)).sum(-1)
    np.fill_diagonal(D, np.inf)
    nbr = np.argsort(D, 1, kind="stable")[:, :k]
    qa, qb = np.repeat(nodes, k, 0), nodes[nbr.ravel()]
    alive = np.ones(len(qa), bool)
    for i in range(n_steps):                       # early exit per edge, like the reference loop
        idx = np.nonzero(alive)[0]
        if len(idx) == 0:
            break
        t = i * step
        hit = ck.config_collision(qa[idx] + t * (qb[idx] - qa[idx]), 0)
        alive[idx[hit]] = False
    return len(qa), time.perf_counter() - t0


def _cpu_worker(args):
    return _cpu_sample(*args)


def cpu_baseline(n_nodes=120, k=10, n_steps=50, step=0.02, procs=1, seed=0):
    if procs == 1:
        edges, sec = _cpu_sample(seed, n_nodes, k, n_steps, step)
        return edges / sec, edges, sec
    import multiprocessing as mp
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(seed + i, n_nodes, k, n_steps, step) for i in range(procs)])
    sec = time.perf_counter() - t0
    edges = sum(r[0] for r in res)
    return edges / sec, edges, sec


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the reference is pure Python
    + pybullet, which is not installable) on all host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    cores = os.cpu_count() or 1
    n_nodes = args.cpu_nodes
    for _ in range(args.warmup and 1):
        cpu_baseline(max(20, n_nodes // 4), args.k, procs=cores)
    t0 = time.perf_counter()
    edges = 0
    for s in range(args.steps):
        _, e, _ = cpu_baseline(n_nodes, args.k, procs=cores, seed=100 + s * cores)
        edges += e
    sec = time.perf_counter() - t0
    value = edges / sec
    sample = (f"{cores} processes x {n_nodes} nodes x k={args.k} x 50 steps per step, one Panda of the 4-arm scene, "
              "hull-GJK checker + float64 FK-metric k-NN")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * sec / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args):
    return {"workload": f"C5: {args.arms} Franka Pandas (circle r=0.665 m, ground plane), {args.nodes} free "
                        f"joint-space samples per arm, k={args.k}, FK metric, 50 steps/edge (local_step 0.02)",
            "arms": args.arms, "nodes_per_arm": args.nodes, "k": args.k, "steps_per_edge": 50,
            "metric_space": "panda task_space_pose_distance (15-D reduced FK features)",
            "l2": "per-step working set (>1 GB) exceeds the 126 MB L2 and a 256 MB buffer is rewritten between steps"}


def ncu_traffic(path):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the scan kernel, from the committed
    summary of the round's `ncu --set full` capture (tools/summarise_profiles.py); None if it is missing."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, seen = 0.0, 0
    try:
        with open(path) as f:
            for row in f:
                cells = row.strip().split(",")
                if len(cells) >= 3 and cells[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(cells[2]) * scale[cells[1]]
                    seen += 1
    except (OSError, ValueError, KeyError):
        return None
    return int(total) if seen == 2 else None


def kernel_counters(core, builder, args):
    """Library counters of one arm's scan + edge check (untimed diagnostic launch)."""
    import ctypes as C
    from mmmp_b200 import _lib
    try:
        L = _lib.load()
        q64, q32 = builder.sample_free(args.nodes, seed=4242)
        L.mmmp_knn_stats(1, None)
        L.mmmp_edge_stats(1, None)
        res = builder.candidates(q64, q32, gather=False)
        lo, hi = res["rows"]
        ks, es = (C.c_ulonglong * 5)(), (C.c_ulonglong * 4)()
        L.mmmp_knn_stats(1, ks)
        L.mmmp_edge_stats(1, es)
        L.mmmp_knn_stats(0, None)
        L.mmmp_edge_stats(0, None)
        nq = max(hi - lo, 1)
        ctas = -(-nq // 128)
        return {"tiles": int(ks[3]), "tiles_per_cta": ks[3] / ctas, "exact_per_query": ks[0] / nq,
                "insert_per_query": ks[1] / nq, "evals_per_edge": (es[1] + es[3]) / max(es[0], 1),
                "deferred_frac": es[2] / max(es[0], 1)}
    except Exception as exc:  # diagnostics must never cost the bench line
        sys.stderr.write(f"kernel_counters failed: {exc}\n")
        return None


# ------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mmmp_b200 import _build, core
    from mmmp_b200.roadmap import PhaseTimer, RoadmapBuilder

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the roadmap path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))   # a lost rank fails fast
    if rank == 0:
        _build.build()
    if world_size > 1:
        dist.barrier()
    dev = torch.device("cuda", local_rank)

    bases, obstacles = scene(args.arms)
    model = core.load_model("panda_spheres")
    world = core.SpatialWorld([core.RobotSpec(model, b, 7 * i) for i, b in enumerate(bases)], obstacles)
    timer = PhaseTimer()
    builders = [RoadmapBuilder(world, r, "fk", args.k, 50, 0.02, timer=None) for r in range(args.arms)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(s, keep=False):
        outs = []
        for r, b in enumerate(builders):
            out = b.build(args.nodes, seed=1000 * (s + 1) + r)
            outs.append(out if keep else None)
        flush.fill_(s & 0xFF)
        return outs

    for s in range(args.warmup):
        step(s)
    # timed region: K steps, device resident
    for b in builders:
        b.timer = timer
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    launches0 = core.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(args.steps):
        timer.mark(None)
        step(args.warmup + s)
    ev1.record()
    barrier()
    launches = core.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    phases = timer.collect()
    clk = clocks.stop() if rank == 0 else None
    for b in builders:
        b.timer = None
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
    if world_size > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, launches = float(tmax[0]), int(tsum[1])
    edges_per_step = args.arms * args.nodes * args.k
    value = edges_per_step * args.steps / (ms / 1000.0)

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region ----
    outs = step(args.warmup + args.steps, keep=True)          # nodes to feed the host-side call
    host_q = []
    for o in outs:
        pin = torch.empty(o["q64"].shape, dtype=torch.float64, pin_memory=True)
        pin.copy_(o["q64"])
        host_q.append(pin)
    del outs
    torch.cuda.synchronize()
    h2d = sum(p.numel() * 8 for p in host_q)
    d2h = args.arms * args.nodes * (4 + args.k * (4 + 8 + 1))

    def e2e_step():
        got = 0
        for r in range(args.arms):
            if world_size == 1:
                res = world.roadmap_candidates_host(host_q[r].numpy(), r, core.METRIC_GROUP_L2_SUM, args.k, 50, 0.02,
                                                    pinned=True)
                got += int(res["n_nodes"])
            else:
                q64 = host_q[r].to(dev, non_blocking=True)
                q32 = core.pack_f32(q64)
                hit = world.collide_configs(q32, r)            # the accept test on the caller's samples
                res = builders[r].candidates(q64, q32)
                if rank == 0:
                    host = [res[key].cpu() for key in ("idx", "dist", "edge_free")] + [hit.cpu()]
                    got += host[0].shape[0]
        return got

    for _ in range(min(args.warmup, 1)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = max(e0.elapsed_time(e1), e2e_wall * 1000.0)      # host-side call: wall clock dominates
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = edges_per_step * args.steps / (float(te[0]) / 1000.0)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        # Two kernels share a step about equally: the neighbour scan and the
        # edge-collision pair.  `roofline` is the scan, reported against HBM as north_star asks, with
        # the ALGORITHMIC streamed bytes of SURVEY.md 8d for an exhaustive scan,
        #     ceil(Nq / Tq) * Nr * Dp * 4 + Nq * kc * 8,
        # Tq = 128 queries per CTA, Dp = 16 (15-D FK features padded), kc = k + 2 candidates.  The kernel
        # skips every tile its bounding-box test proves irrelevant and its rows stay in the 126 MB L2,
        # so the fraction can exceed 1: `executed` carries what one launch really moved and evaluated
        # (library counters, read in an untimed launch after the timed region).
        n_launch = args.steps * args.arms
        scan_ms = phases.get("knn_scan", 0.0) / max(n_launch, 1)
        edge_ms = phases.get("edge_collision", 0.0) / max(n_launch, 1)
        nq = (args.nodes + world_size - 1) // world_size
        TQ, DP, kc = 128, 16, args.k + 2
        alg_bytes = -(-nq // TQ) * args.nodes * DP * 4 + nq * kc * 8
        pairs = nq * args.nodes
        scan_flops = pairs * (5 * 9 + 5)             # FK-metric pair = 5 groups x 9 flop + 5
        sm_count, clock_mhz = 148, float(peaks.get("sm_max_mhz", 1965.0))
        fp32_peak = sm_count * 128 * 2 * clock_mhz * 1e6 / 1e12
        # rank 0 alone, on a builder that does no collectives
        counters = kernel_counters(core, RoadmapBuilder(world, 0, "fk", args.k, 50, 0.02, single_rank=True), args)
        if counters and world_size > 1:
            counters["tiles"] //= world_size      # a rank scans 1/world_size of the query blocks
        tile_bytes = 128 * (4 + 15) * 4              # filter row (4 floats) + exact row (15 floats) per reference
        executed = None
        if counters:
            executed = {"tiles_loaded_per_cta": counters["tiles_per_cta"], "tiles_total_per_cta": -(-args.nodes // 128),
                        "bytes_through_tma_per_launch": counters["tiles"] * tile_bytes,
                        "exact_evaluations_per_query": counters["exact_per_query"],
                        "insertions_per_query": counters["insert_per_query"]}
        # edge pair: per configuration 504 flop of FK + 7 self link-pair and 10
        # link-vs-plane bounding tests (8 and 4 flop) = 600 flop before any sphere-level descent;
        # S = 50 steps per edge, S_eff = steps actually evaluated (library counter)
        edge_cfgs = nq * args.k * 50
        cfg_flop = 504 + 7 * 8 + 10 * 4
        s_eff = counters["evals_per_edge"] if counters else None
        roof = {"bound": "hbm", "kernel": "knn_scan_kernel<1,1>", "achieved": alg_bytes / (scan_ms * 1e-3) / 1e9,
                "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / (scan_ms * 1e-3) / 1e9 / hbm_peak,
                "traffic": ncu_traffic(os.path.join(ROOT, "profiles", "r01_knn_scan_full.csv")),
                "traffic_source": "ncu --set full, profiles/r01_knn_scan_full.csv "
                                                        "(dram__bytes_read.sum + dram__bytes_write.sum, one launch)",
                "peak_source": peak_src, "ms_per_launch": scan_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "executed": executed,
                "binding": "shared-memory/FP32 issue: the packed rows (76 MB per arm) stay in the 126 MB L2 and "
                           "tile culling loads ~3 % of the tiles, so HBM is not the limiter; issue slots are "
                           "(profiles/r01_knn_scan_full.csv: issue active, stall reasons)",
                "fp32": {"achieved": scan_flops / (scan_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": scan_flops / (scan_ms * 1e-3) / 1e12 / fp32_peak,
                         "note": "exhaustive-equivalent flops (SURVEY 8d); culling and the 3-flop centroid filter "
                                 "execute a small fraction of them",
                         "peak_source": "148 SMs x 128 lanes x 2 x clocks.max.sm"},
                "second_kernel": {"kernel": "collide_edges_adaptive_kernel<4> + collide_edges_hard_kernel",
                                  "ms_per_launch_pair": edge_ms, "steps_per_edge": 50, "steps_evaluated_per_edge": s_eff,
                                  "edges_deferred_to_warp_pass": counters["deferred_frac"] if counters else None,
                                  "configs_per_s_nominal": edge_cfgs / (edge_ms * 1e-3) if edge_ms > 0 else None,
                                  "configs_per_s_evaluated": (edge_cfgs * s_eff / 50) / (edge_ms * 1e-3)
                                  if edge_ms > 0 and s_eff else None,
                                  "tflops_evaluated_lower_bound": (edge_cfgs * s_eff / 50) * cfg_flop / (edge_ms * 1e-3) / 1e12
                                  if edge_ms > 0 and s_eff else None,
                                  "fp32_peak": fp32_peak, "bound": "FP32 pipe / issue (divergent sphere-level descents)"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "roadmap_build_ms": ms / args.steps, "configs_checked_per_s_upper": value * 50,
            "phases_ms_per_step": {k_: v / args.steps for k_, v in phases.items()},
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(te[0]) / args.steps,
                    "api": "mmmp_roadmap_candidates_host (C ABI, pinned host buffers)" if world_size == 1 else
                           "RoadmapBuilder.candidates on pinned host samples, results gathered to rank 0 and copied back"},
            "roofline": roof,
        }
        if world_size == 1 and not args.no_cpu:
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
            v, e, sec = cpu_baseline(args.cpu_nodes, args.k)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{args.cpu_nodes} nodes x k={args.k} x 50 steps of one Panda of the same scene "
                                              f"({e} edges in {sec:.1f} s): hull-GJK checker + float64 FK-metric k-NN"}
        print(json.dumps(line))
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--arms", type=int, default=4)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--cpu-nodes", type=int, default=120)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
