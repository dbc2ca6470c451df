#!/usr/bin/env python
"""Micro-benchmark of the neighbour-scan kernel (GPU box): pairs/s and rare-path counters
for the FK metric and joint-space L2 at several queries-per-thread settings.

    python tools/knn_probe.py [n_points]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mmmp_b200 import _lib, core  # noqa: E402

model = core.load_model("panda_spheres")
w = core.SpatialWorld([core.RobotSpec(model, core.base_matrix((0.665, 0, 0.01), (0, 0, 1, 0)))], [core.plane()])
lim = np.array(model["joint_limits"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
q64, q32 = core.sample_uniform(lim[:, 0], lim[:, 1], n, 7)
L = _lib.load()
for name, F in (("fk", w.fk_features(q64, q32, 0)), ("l2", core.l2_features(q64, q32))):
    for Q in (1, 2):
        os.environ["MMMP_KNN_Q"] = str(Q)
        L.mmmp_knn_stats(1, None)
        for it in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ci, cd = core.knn(F, F, 16)
            e1.record()
            torch.cuda.synchronize()
        st = (C.c_ulonglong * 5)()
        L.mmmp_knn_stats(1, st)
        ms = e0.elapsed_time(e1)
        print(f"{name} n={n} Q={Q}: {ms:.1f} ms  {n * n / ms / 1e6:.1f} Gpairs/s  exact/query={st[0] / 2 / n:.0f} "
              f"ins/query={st[1] / 2 / n:.0f} drains/warp={st[2] / 2 / (n / 32 / Q):.0f} "
              f"tiles loaded/CTA={st[3] / 2 / max(n / 128 / Q, 1):.0f} (max {st[4]}) of {(n + 127) // 128}", flush=True)
        L.mmmp_knn_stats(0, None)
