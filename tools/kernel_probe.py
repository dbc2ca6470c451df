#!/usr/bin/env python
"""One launch of one kernel of the path at the bench size, for ncu (profile_r02.sh) -- and, run
without ncu, the library's own work counters of that launch as a line `COUNTERS {json}`:

    python tools/kernel_probe.py scan|scan14|edge|knnedge|configs|fk|accept [n]

scan    knn_scan_kernel<1,1>: 1 M x 1 M FK-metric self search, 12 candidates (the bench launch)
scan14  knn_scan_kernel<4,1>: 200 k composite 14-D rows, L2, 12 candidates (C4)
edge    collide_edges_adaptive_kernel<4> + collide_edges_hard_kernel on the 10 M candidate edges
knnedge mmmp_collide_knn_edges on the same table: mutual candidates evaluated once (the bench's call)
configs collide_configs_kernel on 2 M uniform samples (the accept test of the sampler)
fk      fk_chain_kernel on 1 M configurations (positions only)
accept  the greedy accept pass + components on the candidate tables of one arm
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from mmmp_b200 import _lib, core  # noqa: E402
from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF  # noqa: E402
from mmmp_b200.roadmap import RoadmapBuilder  # noqa: E402

what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else (200_000 if what == "scan14" else 1_000_000)
L = _lib.load()
model = core.load_model("panda_spheres")
bases, obstacles = bench.scene(4)
F = CHECK_SELF | CHECK_OBSTACLES


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)


counters = {"what": what, "n": n}
if what in ("scan", "edge", "knnedge", "accept", "configs", "fk"):
    world = core.SpatialWorld([core.RobotSpec(model, bases[0], 0)], obstacles)
    b = RoadmapBuilder(world, 0, "fk", 10, 50, 0.02, single_rank=True)
    if what == "configs":
        lim = np.array(model["joint_limits"])
        q64, q32 = core.sample_uniform(lim[:, 0], lim[:, 1], 2 * n, seed=1)
        hit, ms = timed(lambda: world.collide_configs(q32, 0, F))
        counters.update(ms=ms, configs=2 * n, colliding=float(hit.float().mean()))
    elif what == "fk":
        q64, q32 = b.sample_free(n, seed=1000)
        _, ms = timed(lambda: world.fk(q32, 0))
        counters.update(ms=ms, configs=n)
    else:
        q64, q32 = b.sample_free(n, seed=1000)
        feats = b.features(q64, q32)
        if what == "scan":
            L.mmmp_knn_stats(1, None)
            (ci, cd), ms = timed(lambda: core.knn(feats, feats, 12))
            ks = (C.c_ulonglong * 5)()
            L.mmmp_knn_stats(1, ks)
            L.mmmp_knn_stats(0, None)
            counters.update(ms=ms, exact_evals=int(ks[0]), insertions=int(ks[1]), drains=int(ks[2]), tile_visits=int(ks[3]),
                            max_tiles_one_cta=int(ks[4]), ctas=-(-n // 128))
        else:
            ci, cd = core.knn(feats, feats, 12)
            idx, dist, amb = core.knn_refine(feats.x64, feats.x64, feats.m64, ci, cd, 10)
            edges = core.edges_from_knn(idx)
            if what == "edge":
                L.mmmp_edge_stats(1, None)
                hit, ms = timed(lambda: world.collide_edges(q32, edges, 50, 0.02, 0, F))
                es = (C.c_ulonglong * 4)()
                L.mmmp_edge_stats(1, es)
                L.mmmp_edge_stats(0, None)
                counters.update(ms=ms, edges=int(es[0]), evals_adaptive=int(es[1]), deferred=int(es[2]), evals_hard=int(es[3]),
                                blocked=float(hit.float().mean()))
            elif what == "knnedge":   # the bench's call: mutual candidates evaluated once
                world.collide_knn_edges(q32, idx, 50, 0.02, 0, F)
                L.mmmp_edge_stats(1, None)
                free, ms = timed(lambda: world.collide_knn_edges(q32, idx, 50, 0.02, 0, F))
                es = (C.c_ulonglong * 4)()
                L.mmmp_edge_stats(1, es)
                L.mmmp_edge_stats(0, None)
                counters.update(ms=ms, slots=int(idx.numel()), edges=int(es[0]), evals_adaptive=int(es[1]), deferred=int(es[2]),
                                evals_hard=int(es[3]), blocked=1.0 - float(free.float().mean()))
            else:
                free = world.collide_knn_edges(q32, idx, 50, 0.02, 0, F)
                (adj, deg, rounds), ms = timed(lambda: core.greedy_degree_connect_device(idx, free, 10, return_rounds=True))
                (labels, count), ms2 = timed(lambda: core.connected_components(adj))
                counters.update(ms=ms, rounds=rounds, components_ms=ms2, mean_degree=float(deg.float().mean()),
                                n_components=count)
elif what == "scan14":
    b2 = [core.base_matrix((0.7, 0, 0.01), (0, 0, 1, 0)), core.base_matrix((-0.7, 0, 0.01))]
    world = core.SpatialWorld([core.RobotSpec(model, x, 7 * i) for i, x in enumerate(b2)], obstacles)
    b = RoadmapBuilder(world, -1, "l2", 10, 20, 0.05, flags=F | CHECK_ROBOTS, single_rank=True)
    q64, q32 = b.sample_free(n, seed=5)
    feats = b.features(q64, q32)
    L.mmmp_knn_stats(1, None)
    (ci, cd), ms = timed(lambda: core.knn(feats, feats, 12))
    ks = (C.c_ulonglong * 5)()
    L.mmmp_knn_stats(1, ks)
    L.mmmp_knn_stats(0, None)
    counters.update(ms=ms, exact_evals=int(ks[0]), insertions=int(ks[1]), tile_visits=int(ks[3]), ctas=-(-n // 128))
else:
    raise SystemExit(__doc__)
print("COUNTERS " + json.dumps(counters), flush=True)
