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
              "hull-GJK checker + float64 FK-metric k-NN; NOTE a 120-node roadmap is far sparser than the 1 M-node "
              "one, so its k-NN edges are much longer (more of them blocked, fewer steps before the early exit)")
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
                        f"joint-space samples per arm, k1={args.k} candidates under the FK metric, 50 steps/edge "
                        f"(local_step 0.02), degree cap k2={args.k2}; one step = samples -> candidate edges checked -> "
                        "accepted adjacency + components of every arm, adjacency handed to rank 0",
            "arms": args.arms, "nodes_per_arm": args.nodes, "k": args.k, "k2": args.k2, "steps_per_edge": 50,
            "metric_space": "panda task_space_pose_distance (15-D reduced FK features)",
            "l2": "per-step working set (>1 GB) exceeds the 126 MB L2 and a 256 MB buffer is rewritten between steps"}


def load_json(path, default=None):
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return default


def kernel_counters(core, builder, args):
    """Library counters of one arm's scan + edge check (untimed diagnostic launch, rank 0 alone)."""
    import ctypes as C
    from mmmp_b200 import _lib
    try:
        L = _lib.load()
        q64, q32 = builder.sample_free(args.nodes, seed=4242)
        L.mmmp_knn_stats(1, None)
        L.mmmp_edge_stats(1, None)
        builder.candidates(q64, q32)
        ks, es = (C.c_ulonglong * 5)(), (C.c_ulonglong * 4)()
        L.mmmp_knn_stats(1, ks)
        L.mmmp_edge_stats(1, es)
        L.mmmp_knn_stats(0, None)
        L.mmmp_edge_stats(0, None)
        nq = max(args.nodes, 1)
        return {"tile_visits": int(ks[3]), "tiles_per_cta": ks[3] / -(-nq // 128), "exact_per_query": ks[0] / nq,
                "insert_per_query": ks[1] / nq, "edges": int(es[0]), "evals_adaptive": int(es[1]),
                "evals_hard": int(es[3]), "evals_per_edge": (es[1] + es[3]) / max(es[0], 1),
                "deferred_frac": es[2] / max(es[0], 1)}
    except Exception as exc:  # diagnostics must never cost the bench line
        sys.stderr.write(f"kernel_counters failed: {exc}\n")
        return None


def limiter(all_phases, teams):
    """Names what holds a multi-GPU step up: the rank with the longest step, its largest phases, and
    the time the ranks of a team spend on work that does not shrink with the team (replicated on
    every rank of it or spent in the exchanges)."""
    totals = [sum(p.values()) for p in all_phases]
    slow = int(np.argmax(totals))
    top = sorted(all_phases[slow].items(), key=lambda kv: -kv[1])[:3]
    fixed = ("sample", "features", "accept_pass", "components", "gather_samples", "gather_adjacency", "handover_to_rank0")
    return {"slowest_rank": slow, "its_ms": round(totals[slow], 3), "its_largest_phases": top,
            "team_size": len(teams.ranks_of_arm[0]),
            "ms_not_shrinking_with_the_team": round(sum(all_phases[slow].get(k_, 0.0) for k_ in fixed), 3),
            "spread_between_ranks_ms": round(max(totals) - min(totals), 3)}


def csv_row(learning_time_s, n_nodes, deg_sum, max_edge_len=None):
    """One row of the reference's experiment CSV schema (experiments/pybullet/exp_scalability.py:300-350)."""
    n_edges = int(deg_sum) // 2
    return {"learning_time": learning_time_s, "n_nodes": int(n_nodes), "n_edges": n_edges,
            "avg_degree": float(deg_sum) / max(int(n_nodes), 1), "max_edge_len": max_edge_len}


def small_configs(core, model):
    """BASELINE.json configs[0..3] on one GPU, one build each after a warm-up build, as rows of the
    reference's CSV schema (learning_time in seconds, CUDA-event timed, device resident)."""
    import torch
    from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF
    from mmmp_b200.roadmap import RoadmapBuilder
    rows = {}

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = fn()
        e1.record()
        torch.cuda.synchronize()
        return res, e0.elapsed_time(e1) / 1000.0

    try:
        # C1: 2 planar 2-link arms (experiments/planar/exp_scalability.py:39-48), 1 k samples per arm,
        # Distance-PRM radius 0.36 (mean degree ~ 10), S = 20, through the planar planner class itself
        from mmmp_b200.planners.prm.planar.env import Environment as PlanarEnv
        from mmmp_b200.planners.prm.planar.multi_arm.decoupled.cbs_prm import CBSPRM as PlanarCBSPRM
        from mmmp_b200.robots.planar_arm import PlanarArm
        arms = [PlanarArm(np.array([3.0, 3.0]), np.array([-5.6, -4.0])), PlanarArm(np.array([3.0, 3.0]), np.array([5.6, -4.0]))]
        agents = [{"name": "a0", "start": np.array([0.0, 0.0]), "goal": np.array([np.pi / 3, 0.0]), "model": arms[0]},
                  {"name": "a1", "start": np.array([np.pi / 3 * 2, 0.0]), "goal": np.array([np.pi, 0.0]), "model": arms[1]}]
        env = PlanarEnv((20, 20), agents, [])
        t0 = time.perf_counter()
        prm = PlanarCBSPRM(env, max_edge_len=0.36, build_type="n_nodes", n_nodes=1000, seed=1)
        wall = time.perf_counter() - t0
        rows["C1"] = {"learning_time": wall, "n_nodes": prm.compute_combined_nr_nodes(),
                      "n_edges": prm.compute_combined_nr_edges(), "avg_degree": prm.compute_combined_avg_degree(),
                      "max_edge_len": 0.36, "note": "planner constructor wall time (includes the Python adjacency lists), "
                      "2 planar 2-link arms x 1000 samples, radius PRM"}
    except Exception as exc:
        rows["C1"] = {"error": repr(exc)}
    try:
        plane = core.plane((0, 0, 1), 0.0)
        boxes = [core.box((0.45, 0.1, 0.25), (0.05, 0.2, 0.25)), core.box((-0.3, -0.45, 0.5), (0.1, 0.1, 0.1)),
                 core.box((0.0, 0.55, 0.3), (0.2, 0.05, 0.3)), core.box((0.5, -0.4, 0.1), (0.1, 0.1, 0.1))]
        w2 = core.SpatialWorld([core.RobotSpec(model, core.base_matrix((0, 0, 0.01)), 0)], [plane] + boxes)
        b2 = RoadmapBuilder(w2, 0, "fk", 10, 50, 0.02, single_rank=True)
        out, sec = timed(lambda: b2.build(10000, seed=2, k2=10))
        rows["C2"] = csv_row(sec, 10000, int(out["deg"].sum()))
        bases = [core.base_matrix((0.7, 0, 0.01), (0, 0, 1, 0)), core.base_matrix((-0.7, 0, 0.01))]
        w3 = core.SpatialWorld([core.RobotSpec(model, b, 7 * i) for i, b in enumerate(bases)], [plane])
        b3 = [RoadmapBuilder(w3, r, "fk", 10, 50, 0.02, single_rank=True) for r in range(2)]
        outs, sec = timed(lambda: [b.build(50000, seed=3 + r, k2=10) for r, b in enumerate(b3)])
        rows["C3"] = csv_row(sec, 100000, sum(int(o["deg"].sum()) for o in outs))
        F = CHECK_SELF | CHECK_OBSTACLES | CHECK_ROBOTS
        b4 = RoadmapBuilder(w3, -1, "l2", 10, 20, 0.05, flags=F, single_rank=True)
        out, sec = timed(lambda: b4.build(200000, seed=5, k2=10))
        rows["C4"] = csv_row(sec, 200000, int(out["deg"].sum()))
        # radius search at C2 size (connect_nearest_neighbors_distance, decoupled_prm.py:553-571)
        q64, q32 = b2.sample_free(10000, seed=2)
        feats = b2.features(q64, q32)
        (off, _, _), sec = timed(lambda: core.radius_search(feats, feats, 0.6, strict=True, exclude_zero=True))
        rows["radius_C2"] = {"learning_time": sec, "n_nodes": 10000, "pairs_within_r": int(off[-1]), "max_edge_len": 0.6,
                             "note": "mmmp_radius alone, FK metric, two-pass CSR"}
    except Exception as exc:
        rows["error"] = repr(exc)
    return rows


# ------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mmmp_b200 import _build, core
    from mmmp_b200.roadmap import ArmTeams, PhaseTimer, RoadmapBuilder

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
    # Arm-parallel layout: every arm's roadmap is independent, so with N <= arms ranks a
    # rank builds whole arms alone (no data-path collective); with N = 8 two ranks share an arm.
    teams = ArmTeams(args.arms)
    timer = PhaseTimer()
    builders = {}
    for a in teams.my_arms:
        group, single = teams.team(a)
        builders[a] = RoadmapBuilder(world, a, "fk", args.k, 50, 0.02, group=group, single_rank=single)
    team_size = len(teams.ranks_of_arm[0])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    like = {a: [((args.nodes, args.k2), torch.int32), ((args.nodes,), torch.int32)] for a in range(args.arms)}

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def seed_of(s, a):
        return 1000 * (s + 1) + a

    def step(s):
        outs = {a: builders[a].build(args.nodes, seed=seed_of(s, a), k2=args.k2) for a in teams.my_arms}
        got = teams.collect_on_root({a: [o["adj"], o["deg"]] for a, o in outs.items()}, like)
        if timer is builders[teams.my_arms[0]].timer:
            timer.mark("handover_to_rank0")
        flush.fill_(s & 0xFF)
        return outs, got

    outs = got = None
    for s in range(args.warmup):
        outs, got = step(s)
    # ---- shard parity (untimed): what the ranks built together equals a single-rank build ----
    shard_parity = None
    if world_size > 1 and args.warmup > 0:
        ok = True
        if rank == 0:
            s = args.warmup - 1
            remote = next((a for a in range(args.arms) if teams.leader(a) != 0), None)
            if remote is not None:      # an arm another rank (team) built and handed over
                ref = RoadmapBuilder(world, remote, "fk", args.k, 50, 0.02, single_rank=True).build(
                    args.nodes, seed=seed_of(s, remote), k2=args.k2)
                ok = ok and torch.equal(ref["adj"], got[remote][0]) and torch.equal(ref["deg"], got[remote][1])
            if team_size > 1:       # arm 0 was sharded over this rank's team
                ref = RoadmapBuilder(world, 0, "fk", args.k, 50, 0.02, single_rank=True).build(
                    args.nodes, seed=seed_of(s, 0), k2=args.k2)
                ok = ok and all(torch.equal(ref[key], outs[0][key]) for key in ("q64", "idx", "dist", "edge_free", "adj"))
            del ref
        shard_parity = bool(ok)
    del outs, got
    # timed region: K steps, device resident
    for b in builders.values():
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
    all_phases = [phases]
    if world_size > 1:          # every rank's phase table, to name the limiter of a multi-GPU step
        all_phases = [None] * world_size
        dist.all_gather_object(all_phases, {k_: round(v / args.steps, 3) for k_, v in phases.items()})
    clk = clocks.stop() if rank == 0 else None
    for b in builders.values():
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

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region ----------------
    # Every rank is handed the samples of ITS arms in page-locked host memory and gets its arms'
    # roadmaps back in page-locked host memory.  One rank per arm: the C-ABI call
    # mmmp_roadmap_build_host; a team of ranks per arm: RoadmapBuilder.build_host (each rank moves
    # only its row block over PCIe).
    host_q = {}
    for a in teams.my_arms:
        q64, _ = builders[a].sample_free(args.nodes, seed=seed_of(args.warmup + args.steps, a))
        pin = torch.empty(q64.shape, dtype=torch.float64, pin_memory=True)
        pin.copy_(q64)
        host_q[a] = pin
        del q64
    torch.cuda.synchronize()
    bufs = {a: (core.host_buffers(args.nodes, args.k, args.k2) if team_size == 1 else {}) for a in teams.my_arms}
    h2d = args.arms * args.nodes * 7 * 8
    d2h = args.arms * args.nodes * (4 + args.k * (4 + 8 + 1) + args.k2 * 4 + 4 + 4)

    # (Calling the arms of a rank from concurrent host threads is supported -- tests/test_gpu_adjacency.py --
    # but measured no faster on average and far less repeatable: 156 ms in one run, 211 ms in the next, as
    # four calls grow the stream-ordered pool against each other.  The arms are called one after the other.)
    def e2e_step():
        got = 0
        for a in teams.my_arms:
            if team_size == 1:
                res = world.roadmap_build_host(host_q[a].numpy(), a, core.METRIC_GROUP_L2_SUM, args.k, args.k2, 50, 0.02,
                                               out=bufs[a])
            else:
                res = builders[a].build_host(host_q[a], args.k2, pinned_out=bufs[a])
            got += int(res["n_nodes"])
        return got

    for _ in range(min(args.warmup, 2)):
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
        peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {})
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        sm_count, clock_mhz = 148, float(peaks.get("sm_max_mhz", 1965.0))
        fp32_peak = sm_count * 128 * 2 * clock_mhz * 1e6 / 1e12
        # launches of rank 0 inside the timed region
        n_launch = args.steps * len(teams.my_arms)
        per = {k_: v / max(n_launch, 1) for k_, v in phases.items()}
        scan_ms, edge_ms = per.get("knn_scan", 0.0), per.get("edge_collision", 0.0)
        counters = kernel_counters(core, RoadmapBuilder(world, 0, "fk", args.k, 50, 0.02, single_rank=True), args)
        share = 1.0 / team_size                      # a rank of a team executes this share of an arm's work
        # Executed FP32 work per unit, MEASURED with ncu on the same workload (thread-level
        # 2*FFMA + FADD + FMUL, smsp__sass_thread_inst_executed_op_f*_pred_on, one launch) and
        # divided by the library's own work counters of that launch: profiles/r02_kernel_flops.json
        # (tools/kernel_flops.py).  bench.py multiplies by the LIVE counters and divides by the LIVE
        # CUDA-event time of the phase.
        kf = load_json(os.path.join(ROOT, "profiles", "r02_kernel_flops.json"), {})

        def roof(name, ms_launch, flop, extra):
            ach = flop / (ms_launch * 1e-3) / 1e12 if ms_launch > 0 and flop else None
            r = {"bound": "fp32", "kernel": name, "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                 "frac": ach / fp32_peak if ach else None, "ms_per_launch": ms_launch,
                 "executed_fp32_flop_per_launch": flop,
                 "peak_source": "148 SMs x 128 FP32 lanes x 2 x clocks.max.sm (MEASURED_PEAKS.json sm_max_mhz)"}
            r.update(extra)
            return r

        edge_roof = scan_roof = None
        if counters:
            e = kf.get("edge_pair", {})
            flop_edge = (counters["evals_adaptive"] * e.get("fp32_flop_per_eval_adaptive", 0.0) +
                         counters["evals_hard"] * e.get("fp32_flop_per_eval_hard", 0.0)) * share
            edge_roof = roof("collide_edges_adaptive_kernel<4> + collide_edges_hard_kernel", edge_ms, flop_edge, {
                "registers": [e.get("adaptive", {}).get("registers"), e.get("hard", {}).get("registers")],
                "traffic": e.get("dram_bytes_per_launch_pair"),
                "unit_of_work": "configuration evaluated (FK of 7 joints + plane / self-pair bounding tests + clearance)",
                "evaluations_per_edge": counters["evals_per_edge"], "steps_per_edge": 50,
                # mutual candidates (i names j and j names i) are evaluated once, low id -> high id
                "candidate_slots": args.nodes * args.k, "pairs_evaluated": counters["edges"],
                "pairs_evaluated_per_slot": counters["edges"] / max(args.nodes * args.k, 1),
                "edges_deferred_to_warp_pass": counters["deferred_frac"],
                "fp32_flop_per_eval": {"adaptive": e.get("fp32_flop_per_eval_adaptive"), "hard": e.get("fp32_flop_per_eval_hard")},
                "pipe_fma_cycles_active_pct": e.get("pipe_fma_cycles_active_pct"),
                "issue_active_pct": e.get("issue_active_pct"), "source": e.get("source"),
                "configs_per_s_nominal": args.nodes * share * args.k * 50 / (edge_ms * 1e-3) if edge_ms > 0 else None,
                "configs_per_s_evaluated": (counters["evals_adaptive"] + counters["evals_hard"]) * share / (edge_ms * 1e-3)
                if edge_ms > 0 else None})
            k_ = kf.get("knn_scan", {})
            flop_scan = counters["tile_visits"] * k_.get("fp32_flop_per_tile_visit", 0.0) * share
            traffic = k_.get("dram_bytes_per_launch")
            scan_roof = roof(k_.get("kernel", "knn_scan_kernel").replace("void ", ""), scan_ms, flop_scan, {
                "registers": k_.get("registers"),
                "traffic": traffic,
                "unit_of_work": "tile visit (128 queries x 128 reference rows through the centroid filter + the exact "
                                "evaluations and insertions it triggers)",
                "tile_visits_per_cta": counters["tiles_per_cta"], "tiles_total_per_cta": -(-args.nodes // 128),
                "exact_evaluations_per_query": counters["exact_per_query"],
                "insertions_per_query": counters["insert_per_query"],
                "pipe_fma_cycles_active_pct": k_.get("pipe_fma_cycles_active_pct"),
                "issue_active_pct": k_.get("issue_active_pct"), "source": k_.get("source"),
                # north_star asks for the k-NN's achieved HBM GB/s: real DRAM traffic of the launch (ncu)
                "hbm": {"achieved": traffic * share / (scan_ms * 1e-3) / 1e9 if traffic and scan_ms > 0 else None,
                        "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
                        "frac": traffic * share / (scan_ms * 1e-3) / 1e9 / hbm_peak if traffic and scan_ms > 0 else None,
                        "note": "the packed rows of an arm (76 MB) stay in the 126 MB L2; culling visits ~2 % of "
                                "the tiles, so DRAM is idle by design"}})
        dominant, other = (edge_roof, scan_roof) if edge_ms >= scan_ms else (scan_roof, edge_roof)
        roofline = dict(dominant or {"bound": "fp32", "kernel": None, "achieved": None, "peak": fp32_peak,
                                     "unit": "TFLOP/s", "frac": None, "traffic": None})
        roofline["second_kernel"] = other
        # what an exhaustive search / check would have cost: kept apart from the roofline fraction
        nq = args.nodes * share
        alg_bytes = -(-int(nq) // 128) * args.nodes * 16 * 4 + nq * (args.k + 2) * 8
        algorithmic = {
            "scan_streamed_bytes_exhaustive": alg_bytes,
            "scan_equivalent_gbs": alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else None,
            "scan_equivalent_tflops": nq * args.nodes * 50 / (scan_ms * 1e-3) / 1e12 if scan_ms > 0 else None,
            "edge_steps_nominal_over_evaluated": 50.0 * args.nodes * args.k / max(counters["evals_adaptive"] + counters["evals_hard"], 1)
            if counters else None,
            "note": "SURVEY 8d brute-force-equivalent figures (Tq=128, Dp=16, 50 flop per FK-metric pair); the kernels "
                    "prune exactly, so these are speed-ups of the algorithm, not utilisation"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "roadmap_build_ms": ms / args.steps,
            "roadmap_build_ends_at": "accepted adjacency (greedy degree pass) + connected components of every arm, "
                                     "adjacency and degrees resident on rank 0",
            "layout": {"ranks_of_arm": teams.ranks_of_arm, "team_size": team_size},
            "shard_parity": shard_parity,
            "configs_checked_per_s_upper": value * 50,
            "phases_ms_per_step_rank0": {k_: v / args.steps for k_, v in phases.items()},
            "phases_ms_per_step_all_ranks": all_phases if world_size > 1 else None,
            "limiter": limiter(all_phases, teams) if world_size > 1 else None,
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(te[0]) / args.steps,
                    "api": "mmmp_roadmap_build_host (C ABI, page-locked host buffers), one call per arm on the rank "
                           "that owns the arm" if team_size == 1 else
                           "RoadmapBuilder.build_host: the ranks of a team upload, check and return their row blocks"},
            "roofline": roofline,
            "algorithmic_speedup": algorithmic,
        }
        if not args.no_configs:
            line["configs"] = small_configs(core, model)
        if world_size == 1 and not args.no_cpu:
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
            v, e, sec = cpu_baseline(args.cpu_nodes, args.k)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{args.cpu_nodes} nodes x k={args.k} x 50 steps of one Panda of the same scene "
                                              f"({e} edges in {sec:.1f} s): hull-GJK checker + float64 FK-metric k-NN; "
                                              "a roadmap this sparse has much longer k-NN edges than the 1 M-node one"}
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
    ap.add_argument("--k2", type=int, default=0, help="degree cap of the accept pass (0: = k)")
    ap.add_argument("--cpu-nodes", type=int, default=120)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1-C4 rows")
    args = ap.parse_args()
    if args.k2 <= 0:
        args.k2 = args.k
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
