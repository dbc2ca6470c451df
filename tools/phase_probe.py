#!/usr/bin/env python
"""Per-kernel timings of the roadmap pipeline for one arm of the C5 scene (GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import scene  # noqa: E402
from mmmp_b200 import core  # noqa: E402
from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_SELF  # noqa: E402


def timed(name, fn, reps=3):
    out = None
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
    print(f"{name:34s} gpu {e0.elapsed_time(e1):9.3f} ms   wall {wall:9.3f} ms", flush=True)
    return out


n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
bases, obstacles = scene(4)
model = core.load_model("panda_spheres")
world = core.SpatialWorld([core.RobotSpec(model, b, 7 * i) for i, b in enumerate(bases)], obstacles)
lim = np.array(model["joint_limits"])
F = CHECK_SELF | CHECK_OBSTACLES
m = 2 * n
q64, q32 = timed("sample_uniform 2n", lambda: core.sample_uniform(lim[:, 0], lim[:, 1], m, 1))
hit = timed("collide_configs 2n (self+obst)", lambda: world.collide_configs(q32, 0, F))
timed("collide_configs 2n (self)", lambda: world.collide_configs(q32, 0, CHECK_SELF))
timed("collide_configs 2n (obst)", lambda: world.collide_configs(q32, 0, CHECK_OBSTACLES))
print("collision rate", float(hit.float().mean()))
idx = timed("compact_free", lambda: core.compact_free(hit))
free64 = timed("gather_rows f64", lambda: core.gather_rows(q64, idx))
free64 = free64[:n].contiguous()
free32 = timed("pack_f32", lambda: core.pack_f32(free64))
feats = timed("fk_features (+fk64)", lambda: world.fk_features(free64, free32, 0))
ci, cd = timed("knn scan kc=16", lambda: core.knn(feats, feats, 16))
ridx, rd, amb = timed("knn refine", lambda: core.knn_refine(feats.x64, feats.x64, feats.m64, ci, cd, 10))
edges = timed("edges_from_knn", lambda: core.edges_from_knn(ridx))
eh = timed("collide_edges 50 steps", lambda: world.collide_edges(free32, edges, 50, 0.02, 0, F))
print("edge collision rate", float(eh.float().mean()), "ambiguous", float(amb.float().mean()))
timed("collide_edges (self only)", lambda: world.collide_edges(free32, edges, 50, 0.02, 0, CHECK_SELF))
timed("collide_edges (obst only)", lambda: world.collide_edges(free32, edges, 50, 0.02, 0, CHECK_OBSTACLES))
eh2 = timed("collide_edges exhaustive", lambda: world.collide_edges(free32, edges, 50, 0.02, 0, F | 16))
print("verdict mismatches vs exhaustive", int((eh2 != eh).sum()))
fk = timed("fk_chain", lambda: world.fk(free32, 0))
free_tab = (eh == 0).to(torch.uint8).reshape(ridx.shape)
labels, count = timed("connected_components (1 M x 10)", lambda: core.connected_components(ridx, free_tab))
print("components", count, "largest", int(torch.bincount(labels.long()).max()))
# batched IK: tool positions of the free nodes as targets, tool pointing down, one shared start pose
tip = fk[:, -1, :].double().contiguous()
quat = torch.tensor([[1.0, 0.0, 0.0, 0.0]], dtype=torch.float64, device=tip.device).repeat(tip.shape[0], 1)
q_start = torch.tensor([[0.0, -0.785, 0.0, -2.356, 0.0, 1.571, 0.785]], dtype=torch.float64, device=tip.device)
qi, conv, res = timed("ik_dls 1 M poses (<= 100 iterations)", lambda: world.ik(tip, quat, q_start, 0))
print("ik converged", float(conv.float().mean()))
