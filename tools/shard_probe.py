#!/usr/bin/env python
"""Per-shard cost of the sharded neighbour search on ONE GPU (what each rank of a G-GPU run
would do): time, tiles loaded per query block (mean / max) for every shard of G = 2, 4, 8.

    python tools/shard_probe.py [n] > profiles/r01_shard_probe.txt
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from mmmp_b200 import _lib, core  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
L = _lib.load()
model = core.load_model("panda_spheres")
bases, obstacles = bench.scene(4)
world = core.SpatialWorld([core.RobotSpec(model, bases[0], 0)], obstacles)
lim = np.array(model["joint_limits"])
q64, q32 = core.sample_uniform(lim[:, 0], lim[:, 1], 2 * n, seed=1)
hit = world.collide_configs(q32, 0, 3)
free64 = core.gather_rows(q64, core.compact_free(hit))[:n].contiguous()
feats = world.fk_features(free64, core.pack_f32(free64), 0)
for G in (1, 2, 4, 8):
    line = []
    for s in range(G):
        core.knn_shard(feats, 12, s, G)
        L.mmmp_knn_stats(1, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rows, _, _ = core.knn_shard(feats, 12, s, G)
        e1.record()
        torch.cuda.synchronize()
        st = (C.c_ulonglong * 5)()
        L.mmmp_knn_stats(1, st)
        L.mmmp_knn_stats(0, None)
        blocks = max(1, -(-rows.shape[0] // 128))
        line.append(f"{e0.elapsed_time(e1):.1f} ms ({st[3] / blocks:.0f}/{st[4]} tiles)")
    print(f"G={G}: " + "  ".join(line), flush=True)
