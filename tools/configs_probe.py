#!/usr/bin/env python
"""Roadmap build times of the smaller BASELINE.json configurations on one GPU (the bench line is
C5): C2 single Panda, 10 k nodes, FK metric and L2, box obstacles, S = 50; C3 two Pandas
(sims/cbs_prm.py bases), 50 k nodes per arm, S = 50; C4 two Pandas coupled, 14-D composite
rows, 200 k nodes, L2, S = 20, self + obstacle + robot-robot checks.

    python tools/configs_probe.py > profiles/r01_configs_probe.txt
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from mmmp_b200 import core  # noqa: E402
from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF  # noqa: E402
from mmmp_b200.roadmap import PhaseTimer, RoadmapBuilder  # noqa: E402

model = core.load_model("panda_spheres")
plane = core.plane((0, 0, 1), 0.0)


def run(name, builders, n, reps=3):
    timer = PhaseTimer()
    for b in builders:
        b.build(n, seed=1)                                  # warm-up
        b.timer = timer
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        timer.mark(None)
        for i, b in enumerate(builders):
            out = b.build(n, seed=10 + 7 * r + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ph = {k: round(v / reps, 3) for k, v in timer.collect().items()}
    edges = len(builders) * n * builders[0].k
    print(f"{name}: {ms:.2f} ms per build, {edges / ms * 1e3:.3e} candidate edges/s, "
          f"free edges {float(out['edge_free'].float().mean()):.3f}, phases(ms) {ph}", flush=True)


boxes = [core.box((0.45, 0.1, 0.25), (0.05, 0.2, 0.25)), core.box((-0.3, -0.45, 0.5), (0.1, 0.1, 0.1)),
         core.box((0.0, 0.55, 0.3), (0.2, 0.05, 0.3)), core.box((0.5, -0.4, 0.1), (0.1, 0.1, 0.1))]
w2 = core.SpatialWorld([core.RobotSpec(model, core.base_matrix((0, 0, 0.01)), 0)], [plane] + boxes)
run("C2 1 Panda 10k k=10 FK metric S=50", [RoadmapBuilder(w2, 0, "fk", 10, 50, 0.02)], 10000)
run("C2 1 Panda 10k k=10 L2        S=50", [RoadmapBuilder(w2, 0, "l2", 10, 50, 0.02)], 10000)
bases = [core.base_matrix((0.7, 0, 0.01), (0, 0, 1, 0)), core.base_matrix((-0.7, 0, 0.01))]
w3 = core.SpatialWorld([core.RobotSpec(model, b, 7 * i) for i, b in enumerate(bases)], [plane])
run("C3 2 Pandas 50k/arm k=10 FK metric S=50", [RoadmapBuilder(w3, r, "fk", 10, 50, 0.02) for r in range(2)], 50000)
F = CHECK_SELF | CHECK_OBSTACLES | CHECK_ROBOTS
run("C4 2 Pandas coupled 14-D 200k k=10 L2 S=20", [RoadmapBuilder(w3, -1, "l2", 10, 20, 0.05, flags=F)], 200000)
