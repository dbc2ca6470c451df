#!/usr/bin/env python
"""Which clearance term limits the step-skipping edge kernel?  Needs the diagnostic variant:
    python tools/build_variant.py diag -DMMMP_EDGE_DIAG
    MMMP_B200_LIB=mmmp_b200/lib/variants/libmmmp_b200_diag.so python tools/edge_diag.py [n]
Prints, for the 10 n candidate edges of one C5 arm: the histogram of steps proven per side by an
evaluation, and for the evaluations that proved less than one step the term that limited them
(obstacle term of joint j / self pair)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from mmmp_b200 import _lib, core  # noqa: E402
from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_SELF  # noqa: E402
from mmmp_b200.roadmap import RoadmapBuilder  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
L = C.CDLL(_lib.lib_path())
model = core.load_model("panda_spheres")
bases, obstacles = bench.scene(4)
world = core.SpatialWorld([core.RobotSpec(model, bases[0], 0)], obstacles)
b = RoadmapBuilder(world, 0, "fk", 10, 50, 0.02, single_rank=True)
q64, q32 = b.sample_free(n, seed=1000)
feats = b.features(q64, q32)
ci, cd = core.knn(feats, feats, 12)
idx, dist, amb = core.knn_refine(feats.x64, feats.x64, feats.m64, ci, cd, 10)
edges = core.edges_from_knn(idx)
dq = (q64[idx.long().ravel()] - q64.repeat_interleave(10, 0)).abs()
print("mean |dq| per joint over the candidate edges:", dq.mean(0).cpu().numpy().round(3))
print("mean FK distance of the candidates:", float(dist.mean()))
pairs = [tuple(p) for p in model["self_pairs"]]
for name, flags in (("self+obst", CHECK_SELF | CHECK_OBSTACLES), ("self", CHECK_SELF), ("obst", CHECK_OBSTACLES)):
    L.mmmp_edge_diag(None)
    hit = world.collide_edges(q32, edges, 50, 0.02, 0, flags)
    torch.cuda.synchronize()
    d = (C.c_ulonglong * 64)()
    L.mmmp_edge_diag(d)
    d = np.array(list(d), dtype=np.float64)
    tot = max(d[63], 1)
    print(f"== {name}: {int(tot)} gap evaluations ({tot / edges.shape[0]:.2f} per edge), blocked {float(hit.float().mean()):.4f}")
    print("   steps proven per side (0..15+):", (d[32:48] / tot).round(3).tolist())
    print("   limiting term when < 1 step: obstacle@joint", (d[0:8] / tot).round(4).tolist(),
          " self pair", (d[16:30] / tot).round(4).tolist(), " none/robot", (d[30:32] / tot).round(4).tolist())
