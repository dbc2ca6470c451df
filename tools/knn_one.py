#!/usr/bin/env python
"""One neighbour scan for profiling: python tools/knn_one.py n fk|fkc|l2 Q [k]   (fkc: centroid filter rows)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mmmp_b200 import core  # noqa: E402

n, metric, Q = int(sys.argv[1]), sys.argv[2], sys.argv[3]
k = int(sys.argv[4]) if len(sys.argv) > 4 else 16
model = core.load_model("panda_spheres")
w = core.SpatialWorld([core.RobotSpec(model, core.base_matrix((0.665, 0, 0.01), (0, 0, 1, 0)))], [core.plane()])
lim = np.array(model["joint_limits"])
q64, q32 = core.sample_uniform(lim[:, 0], lim[:, 1], n, 7)
F = (w.fk_features(q64, q32, 0, dual=metric == "fk") if metric in ("fk", "fkc") else core.l2_features(q64, q32))
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ci, cd = core.knn(F, F, k)
    e1.record()
    torch.cuda.synchronize()
    print(f"{metric} n={n} Q={Q}: {e0.elapsed_time(e1):.2f} ms", flush=True)
# order-sensitive checksum of the answer, to compare tuning builds of the library
mix = torch.arange(1, ci.numel() + 1, device=ci.device, dtype=torch.int64).reshape(ci.shape) % 1000003
print("checksum", int((ci.long() * mix).sum()), float(cd.double().sum()), flush=True)
