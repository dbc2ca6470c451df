"""Device-resident roadmap construction pipeline (one arm), optionally sharded over the
ranks of a torch.distributed process group (one process per GPU, NCCL over NVLink).

Reference path replaced (planners/prm/pybullet/decoupled/decoupled_prm.py):
  sample_valid_in_joint_space_random :310-321  -> sample_free()
  connect_nearest_neighbors_degree   :480-504  -> candidates() (+ host greedy replay)
  is_collision_free_sample / _edge   :840-863  -> collision kernels

Sharding: rank g draws and checks a contiguous block of candidate
samples; the collision-free rows are all-gathered in rank order, which reproduces the
single-GPU node order exactly; every rank then owns the contiguous query rows
[g*n/G, (g+1)*n/G) for neighbour search and edge checks against the full gathered
set, and the per-row results are all-gathered.  No other collective is needed.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import core
from ._lib import CHECK_OBSTACLES, CHECK_SELF, METRIC_GROUP_L2_SUM, METRIC_L2


class PhaseTimer:
    """CUDA-event timer on the current stream; phases accumulate across calls."""

    def __init__(self):
        self.events = []
        self.totals = {}

    def mark(self, name: str):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.events.append((name, e))

    def collect(self):
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(self.events[:-1], self.events[1:]):
            if n1 is None:
                continue
            self.totals[n1] = self.totals.get(n1, 0.0) + e0.elapsed_time(e1)
        self.events = []
        return self.totals


def shard_rows(n: int, rank: int, size: int):
    """Contiguous row block [lo, hi) of rank `rank` out of `size` (last blocks may be short)."""
    per = (n + size - 1) // size
    return min(rank * per, n), min((rank + 1) * per, n)


def all_gather_rows(t: torch.Tensor, size: int, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks in rank order; row counts may differ per rank.
    Works on CUDA tensors over NCCL and on CPU tensors over gloo (tests)."""
    if size == 1:
        return t
    import torch.distributed as dist
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    ns = [torch.empty_like(n) for _ in range(size)]
    dist.all_gather(ns, n, group=group)
    ns = [int(x.item()) for x in ns]
    m = max(ns)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([outs[g][: ns[g]] for g in range(size)], 0)


def all_gather_padded(t: torch.Tensor, cap: int, size: int, group=None, fill=0) -> torch.Tensor:
    """Every rank contributes at most `cap` rows; -> [size * cap, ...] with rank g's rows at
    [g * cap, g * cap + its row count) and `fill` in the padding.  One collective, no size
    exchange, no host synchronisation."""
    import torch.distributed as dist
    pad = torch.full((cap,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    if t.is_cuda:       # NCCL: one flat collective
        out = torch.empty((size * cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, pad, group=group)
        return out
    # CPU tensors (gloo, the host-logic tests): the list form; chosen by where the data lives, never
    # by catching a failed collective -- a rank that fails must fail, not issue a second collective
    parts = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat(parts, 0)


def balanced_cuts(cuts, times, damping: float = 0.5):
    """One step of the cut controller of the sharded neighbour search: `cuts` (G + 1 fractions of
    the scan order, 0 .. 1) and the scan time every rank measured for its slice -> new cuts, moved
    towards the widths that would have taken equal time.  Damped, because the cost per query
    block is only locally constant.  Times that are not all positive leave the cuts alone."""
    cuts = np.asarray(cuts, dtype=np.float64)
    times = np.asarray(times, dtype=np.float64)
    if times.shape[0] != cuts.shape[0] - 1 or not np.all(times > 0):
        return cuts
    width = np.diff(cuts)
    want = width / times                         # slice widths that would have taken equal time
    want /= want.sum()
    width = (1.0 - damping) * width + damping * want
    out = np.concatenate([[0.0], np.cumsum(width)])
    out[-1] = 1.0
    return out


def arm_layout(world_size: int, n_arms: int):
    """Which ranks build which arm's roadmap (in decoupled planners every arm's roadmap
    is independent).  -> ranks_of_arm: list of rank lists, one per arm.
      world_size <= n_arms : arm a belongs to rank a % world_size alone -- no data-path collective;
      world_size a multiple of n_arms : world_size / n_arms consecutive ranks share an arm
                             (sharded sampling, neighbour search and edge checks inside the team);
      otherwise            : every arm is sharded over all ranks."""
    if world_size <= n_arms:
        return [[a % world_size] for a in range(n_arms)]
    if world_size % n_arms == 0:
        per = world_size // n_arms
        return [list(range(a * per, (a + 1) * per)) for a in range(n_arms)]
    return [list(range(world_size)) for _ in range(n_arms)]


class ArmTeams:
    """The process groups of an arm-parallel build and the final hand-over of every arm's result to
    rank 0 ("adjacency resident on rank 0").  Every rank must construct it (new_group is
    collective)."""

    def __init__(self, n_arms: int, rank: Optional[int] = None, world_size: Optional[int] = None):
        import torch.distributed as dist
        live = dist.is_available() and dist.is_initialized()
        self.rank = (dist.get_rank() if live else 0) if rank is None else rank
        self.world_size = (dist.get_world_size() if live else 1) if world_size is None else world_size
        self.ranks_of_arm = arm_layout(self.world_size, n_arms)
        self.groups = {}
        made = {}
        for a, ranks in enumerate(self.ranks_of_arm):
            key = tuple(ranks)
            if len(ranks) > 1 and len(ranks) < self.world_size and key not in made:
                made[key] = dist.new_group(list(ranks))           # every rank calls, in the same order
            self.groups[a] = made.get(key)                         # None: alone, or the whole world
        self.my_arms = [a for a, ranks in enumerate(self.ranks_of_arm) if self.rank in ranks]

    def team(self, arm: int):
        """(group, single_rank) arguments of the RoadmapBuilder of `arm` on this rank."""
        ranks = self.ranks_of_arm[arm]
        return self.groups[arm], len(ranks) == 1

    def leader(self, arm: int) -> int:
        return self.ranks_of_arm[arm][0]

    def collect_on_root(self, per_arm: dict, like: dict):
        """per_arm[a] = list of tensors this rank holds for its arm a (identical on every rank of
        the team); like[a] = list of (shape, dtype) of every arm's tensors.  Rank 0 ends up with
        the tensors of ALL arms (returns {arm: [tensors]}); leaders of the other arms send theirs
        point to point, everything in one batch so that no ordering between pairs can dead-lock."""
        import torch.distributed as dist
        if self.world_size == 1:
            return per_arm
        ops, got = [], {}
        for a in range(len(self.ranks_of_arm)):
            src = self.leader(a)
            if src == 0:
                if self.rank == 0:
                    got[a] = per_arm[a]
                continue
            if self.rank == 0:
                dev = next(iter(per_arm.values()))[0].device if per_arm else torch.device("cpu")
                got[a] = [torch.empty(shape, dtype=dtype, device=dev) for shape, dtype in like[a]]
                ops += [dist.P2POp(dist.irecv, t, src) for t in got[a]]
            elif self.rank == src:
                ops += [dist.P2POp(dist.isend, t.contiguous(), 0) for t in per_arm[a]]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return got if self.rank == 0 else per_arm


class RoadmapBuilder:
    def __init__(self, world: core.SpatialWorld, robot: int = 0, metric: str = "fk", k: int = 10,
                 n_steps: int = 50, step: float = 0.02, flags: int = CHECK_SELF | CHECK_OBSTACLES,
                 group=None, timer: Optional[PhaseTimer] = None, knn_pad: int = 2, single_rank: bool = False):
        self.world, self.robot, self.metric, self.k = world, robot, metric, k
        self.n_steps, self.step, self.flags, self.knn_pad = n_steps, step, flags, knn_pad
        self.group = group
        if single_rank:
            self.rank, self.size = 0, 1
        elif group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.rank = torch.distributed.get_rank(group)
            self.size = torch.distributed.get_world_size(group)
        else:
            self.rank, self.size = 0, 1
        self.timer = timer
        if robot < 0:   # composite rows of all robots (coupled planners, coupled_prm.py:88-109): L2 metric
            if metric != "l2":
                raise ValueError("composite roadmaps use the joint-space L2 metric (general_distance)")
            lim = np.concatenate([np.array(r.model["joint_limits"], dtype=np.float64) for r in world.robots])
        else:
            lim = np.array(world.robots[robot].model["joint_limits"], dtype=np.float64)
        self.lo, self.hi = lim[:, 0].copy(), lim[:, 1].copy()
        self.D = lim.shape[0]
        self.accept_rate = None
        self.cuts = np.linspace(0.0, 1.0, self.size + 1)    # slices of the neighbour search's scan order

    def _rebalance(self, my_scan_ms: float):
        """Move the cuts of the scan order towards equal scan times: slices cost differently (the
        tile count per query block varies by ~10 % between regions of the filter space), and the
        all-gather after the scan makes everybody wait for the slowest rank.  Every rank sees the
        same all-gathered times, so every rank computes the same new cuts."""
        if self.size > 4:
            # measured on 8 B200s: a slice's time is dominated by its longest query block, not by
            # its width, and moving the cuts gains nothing (71.0 ms per step against 69.5 ms)
            return
        import torch.distributed as dist
        t = torch.tensor([my_scan_ms], dtype=torch.float64, device="cuda")
        ts = [torch.empty_like(t) for _ in range(self.size)]
        dist.all_gather(ts, t, group=self.group)
        self.cuts = balanced_cuts(self.cuts, np.array([float(x.item()) for x in ts]))

    def _mark(self, name):
        if self.timer is not None:
            self.timer.mark(name)

    # ------------------------------------------------------------------ collectives
    def _all_gather_rows(self, t: torch.Tensor, counts=None) -> torch.Tensor:
        return all_gather_rows(t, self.size, self.group)

    # ------------------------------------------------------------------ sampling
    def sample_free(self, n: int, seed: int, oversample: float = 1.05, first_index: int = 0):
        """n collision-free joint-space samples (ids = acceptance order of the global
        candidate stream), identical on every rank.  -> (q64 [n, D], q32 [n, ld])"""
        rate = self.accept_rate or 0.5
        got64, got32, have = [], [], 0
        first = first_index
        while have < n:
            m = int(np.ceil((n - have) / rate * oversample)) + 64
            m = ((m + self.size - 1) // self.size) * self.size
            per = m // self.size
            q64, q32 = core.sample_uniform(self.lo, self.hi, per, seed, first + self.rank * per)
            self._mark("sample")
            hit = self.world.collide_configs(q32, self.robot, self.flags)
            self._mark("config_collision")
            idx = core.compact_free(hit)
            f64 = core.gather_rows(q64, idx)
            self._mark("compact")
            f64 = self._all_gather_rows(f64)
            rate = max(f64.shape[0] / m, 1e-3)
            got64.append(f64)
            have += f64.shape[0]
            first += m
        self.accept_rate = rate
        q64 = (got64[0] if len(got64) == 1 else torch.cat(got64, 0))[:n].contiguous()
        q32 = core.pack_f32(q64)
        self._mark("gather_samples")
        return q64, q32

    # ------------------------------------------------------------------ neighbour search + edges
    def features(self, q64: torch.Tensor, q32: torch.Tensor) -> core.Features:
        if self.metric == "l2":
            return core.l2_features(q64, q32)
        return self.world.fk_features(q64, q32, self.robot)

    def candidates(self, q64: torch.Tensor, q32: torch.Tensor, gather: bool = True):
        """k nearest candidates of every node + the verdict of every candidate edge.
        -> dict(idx [n,k] int32, dist [n,k] fp64, edge_free [n,k] uint8, ambiguous [n] uint8).
        With several ranks:
          * neighbour search -- every rank answers one slice of the library's SPATIAL scan order
            (core.knn_shard): its query blocks stay as compact as on one GPU, the tile culling
            keeps its selectivity and a rank does 1/size of the work; the answered rows are
            all-gathered and scattered back to the caller's row order;
          * edge checks -- every rank takes a contiguous block of the caller's rows (the sampling
            order: hard and easy edges mixed, so the ranks finish together), verdicts all-gathered.
        gather=False returns this rank's row block [lo, hi) of every table ("rows")."""
        n = q64.shape[0]
        lo, hi = shard_rows(n, self.rank, self.size)
        feats = self.features(q64, q32)
        self._mark("features")
        # the fp32 scan keeps k + knn_pad candidates for the fp64 re-rank; the few rows the re-rank
        # cannot separate from the last candidate (`ambiguous`, ~1e-8 per query at pad 2) are
        # searched again, alone, with k + 16 candidates
        kc, kc_wide = min(self.k + self.knn_pad, 64), min(self.k + 16, 63)
        if self.size == 1:
            rows, mine64 = None, feats.x64
            ci, cd = core.knn(feats, feats, kc, exclude_self=True, causal=False)
        else:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            rows, ci, cd = core.knn_range(feats, kc, self.cuts[self.rank], self.cuts[self.rank + 1], exclude_self=True)
            t1.record()
            mine64 = core.gather_rows(feats.x64, rows)
        self._mark("knn_scan")
        idx, dist, amb = core.knn_refine(feats.x64, mine64, feats.m64, ci, cd, self.k)
        if kc < kc_wide and bool(amb.any()):
            bad = amb.nonzero().flatten()                       # positions in this rank's tables
            ids = bad if rows is None else rows.long()[bad]     # their row ids
            sub = feats.take(ids)
            ci2, cd2 = core.knn(feats, sub, kc_wide + 1, exclude_self=False)   # self comes back at distance 0 ...
            keep = ci2 != ids.to(torch.int32)[:, None]                          # ... and is dropped here
            order = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)[:, :kc_wide]
            ci2, cd2 = torch.gather(ci2, 1, order).contiguous(), torch.gather(cd2, 1, order).contiguous()
            idx2, dist2, amb2 = core.knn_refine(feats.x64, sub.x64, feats.m64, ci2, cd2, self.k)
            idx[bad], dist[bad], amb[bad] = idx2, dist2, amb2
        self._mark("knn_rerank")
        if rows is not None:   # spatial slices -> the caller's row order, on every rank
            # ONE all-gather of packed records (row id, flag, idx, fp64 dist: 128 B per row at k = 10)
            # and ONE scatter kernel; padding records carry row id -1
            widest = max(b - a for a, b in zip(self.cuts[:-1], self.cuts[1:]))
            cap = (int(np.ceil(-(-n // 128) * widest)) + 2) * 128       # rows the widest slice can hold
            rec = core.pack_neighbour_records(rows, idx, dist, amb, cap)
            allrec = all_gather_padded(rec, cap, self.size, self.group)
            idx, dist, amb = core.scatter_neighbour_records(allrec, self.k, n)
            if self.size <= 4:          # the cuts only move up to 4 ranks (see _rebalance)
                t1.synchronize()
                self._rebalance(t0.elapsed_time(t1))
            self._mark("gather_adjacency")
        # a pair of nodes that name each other is evaluated once (mmmp_collide_knn_edges); across ranks
        # the evaluation belongs to one of the two rows and the other slot is filled in after the gather
        across = self.size > 1 and gather
        free = self.world.collide_knn_edges(q32, idx, self.n_steps, self.step, self.robot, self.flags,
                                            rows=(lo, hi), share_across_rows=across)
        self._mark("edge_collision")
        if self.size > 1:
            if gather:   # equal row blocks (the last may be short): rank order = row order
                per = (n + self.size - 1) // self.size
                free = all_gather_padded(free, per, self.size, self.group)[:n].contiguous()
                core.SpatialWorld.resolve_knn_edges(idx, free)
                self._mark("gather_adjacency")
            else:
                idx, dist, amb = idx[lo:hi], dist[lo:hi], amb[lo:hi]
        return {"idx": idx, "dist": dist, "edge_free": free, "ambiguous": amb, "rows": (lo, hi) if not gather else (0, n)}

    def build(self, n: int, seed: int, k2: Optional[int] = None):
        """Learning phase of one arm, device resident: samples -> candidate tables (k nearest +
        edge verdicts); with k2 also the accepted adjacency and its components (adjacency())."""
        q64, q32 = self.sample_free(n, seed)
        out = self.candidates(q64, q32)
        out["q64"], out["q32"] = q64, q32
        if k2 is not None:
            self.adjacency(out, k2)
        return out

    def adjacency(self, out, k2: Optional[int] = None, cap: Optional[int] = None):
        """The roadmap itself: connect_nearest_neighbors_degree's accept pass (decoupled_prm.py:
        496-503: nodes in id order, candidates in (dist, id) order, insert at both ends while both
        hold fewer than k2 edges) on the gathered candidate tables, then its connected components
        (:613-636).  Adds adj [n, cap] int32 (-1 padded, reference list order), deg [n],
        labels [n], n_components ([1] int64, left on the device) to `out`."""
        k2 = self.k if k2 is None else k2
        out["adj"], out["deg"] = core.greedy_degree_connect_device(out["idx"], out["edge_free"], k2, cap)
        self._mark("accept_pass")
        out["labels"], out["n_components"] = core.connected_components(out["adj"], sync=False)
        self._mark("components")
        return out

    def build_host(self, q_host: torch.Tensor, k2: int, pinned_out: Optional[dict] = None):
        """The host-buffer build for a TEAM of ranks (one rank: SpatialWorld.roadmap_build_host, the
        C-ABI call).  q_host [N, D] fp64 page-locked samples, the same on every rank of the team.
        Every rank uploads and accept-tests only ITS row block, the free rows are all-gathered (node
        ids = input order), the team shares neighbour search and edge checks, every rank runs the
        accept pass and copies ITS row block of every table back into page-locked buffers.
        -> dict of host tensors for rows [lo, hi) of the n nodes: node_src, nbr_idx, nbr_dist,
        edge_free, adj, deg, labels; plus n_nodes, rows."""
        N = q_host.shape[0]
        dev = torch.device("cuda", torch.cuda.current_device())
        a, b = shard_rows(N, self.rank, self.size)
        mine = q_host[a:b].to(dev, non_blocking=True)
        m32 = core.pack_f32(mine)
        hit = self.world.collide_configs(m32, self.robot, self.flags)
        keep = core.compact_free(hit)
        src = keep + a
        q64 = all_gather_rows(core.gather_rows(mine, keep), self.size, self.group)
        src = all_gather_rows(src[:, None].contiguous(), self.size, self.group)[:, 0].contiguous()
        q32 = core.pack_f32(q64)
        self._mark("h2d_accept_gather")
        out = self.candidates(q64, q32)
        self.adjacency(out, k2)
        n = q64.shape[0]
        lo, hi = shard_rows(n, self.rank, self.size)
        host = pinned_out if pinned_out is not None else {}
        tables = {"node_src": src, "nbr_idx": out["idx"], "nbr_dist": out["dist"], "edge_free": out["edge_free"],
                  "adj": out["adj"], "deg": out["deg"], "labels": out["labels"]}
        res = {}
        for key, t in tables.items():
            part = t[lo:hi]
            buf = host.get(key)
            if buf is None or buf.shape[0] < hi - lo or buf.shape[1:] != part.shape[1:] or buf.dtype != part.dtype:
                buf = torch.empty((hi - lo,) + tuple(part.shape[1:]), dtype=part.dtype, pin_memory=True)
                host[key] = buf
            buf[: hi - lo].copy_(part, non_blocking=True)
            res[key] = buf[: hi - lo]
        torch.cuda.current_stream().synchronize()
        self._mark("d2h")
        res.update(n_nodes=n, rows=(lo, hi))
        return res

    def components(self, out):
        """Connected components of the collision-free candidate graph of build() / candidates()
        (gathered tables): labels [n] int32 (smallest node id of the component), component count.
        compute_connected_components (decoupled_prm.py:613-636) on the GPU."""
        labels, count = core.connected_components(out["idx"], out["edge_free"])
        self._mark("components")
        return labels, count
