#!/usr/bin/env python
"""Edge-collision kernel probe on the bench scene (one Panda of the 4-arm ring, ground plane):
roadmap edges of n free nodes (k = 10, FK metric), default step-skipping kernel against the
exhaustive one, with the evaluations-per-edge counter.  GPU box:

    python tools/edge_probe.py [n] [refine,refine,...]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from mmmp_b200 import _lib, core  # noqa: E402
from mmmp_b200._lib import CHECK_EXHAUSTIVE, CHECK_OBSTACLES, CHECK_SELF  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
refines = [float(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0.0]
hards = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [30]
L = _lib.load()
model = core.load_model("panda_spheres")
bases, obstacles = bench.scene(4)
world = core.SpatialWorld([core.RobotSpec(model, bases[0], 0)], obstacles)
F = CHECK_SELF | CHECK_OBSTACLES
lim = np.array(model["joint_limits"])
q64, q32 = core.sample_uniform(lim[:, 0], lim[:, 1], 2 * n, seed=1)
hit = world.collide_configs(q32, 0, F)
free64 = core.gather_rows(q64, core.compact_free(hit))[:n].contiguous()
free32 = core.pack_f32(free64)
feats = world.fk_features(free64, free32, 0)
ci, cd = core.knn(feats, feats, 16)
ridx, rd, amb = core.knn_refine(feats.x64, feats.x64, feats.m64, ci, cd, 10)
edges = core.edges_from_knn(ridx)
E = edges.shape[0]


def timed(flags, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = world.collide_edges(free32, edges, 50, 0.02, 0, flags)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return out, best


ref, ms = timed(F | CHECK_EXHAUSTIVE)
print(f"n={free32.shape[0]} E={E} exhaustive: {ms:.1f} ms  blocked {float(ref.float().mean()):.4f}", flush=True)
for name, flags in (("self+obst", F), ("self", CHECK_SELF), ("obst", CHECK_OBSTACLES)):
    for r in [(a, b) for a in refines for b in hards]:
        os.environ["MMMP_EDGE_REFINE"] = str(r[0])
        os.environ["MMMP_EDGE_HARD"] = str(r[1])
        L.mmmp_edge_stats(1, None)
        out, ms = timed(flags, reps=2)
        st = (C.c_ulonglong * 4)()
        L.mmmp_edge_stats(1, st)
        L.mmmp_edge_stats(0, None)
        extra = f" mismatches {int((out != ref).sum())}" if flags == F else ""
        print(f"{name:10s} refine={r[0]:4.2f} hard_after={r[1]:3d}: {ms:7.1f} ms  evaluations/edge {(st[1] + st[3]) / max(st[0], 1):.2f} "
              f"(deferred {st[2] / max(st[0], 1):.3f} of the edges, {st[3] / max(st[0], 1):.2f} evaluations/edge there){extra}", flush=True)
