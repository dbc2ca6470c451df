#!/usr/bin/env python
"""Collision-verdict disagreement of the sphere model (CUDA kernels) against the convex-hull
oracle (restatement of pybullet's checker, parity unpinned) on uniform Panda configurations.
Prints false-negative / false-positive rates per predicate family, with and without the
exact refinements (proven-free pairs, link5/link7 table).  GPU box:

    python tools/disagreement.py [n] [margin_m] > profiles/r02_disagreement.txt

margin_m inflates every hull of the ORACLE (Bullet's collision margin; its URDF importer is
remembered to set 1 mm per body, which cannot be checked here: no pybullet anywhere, see
profiles/r02_pybullet_probe.txt).  With a margin the sphere cover, fitted to the bare hulls, may
miss grazing contacts: the false-negative column then measures how much a 1 mm margin would matter.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from mmmp_b200 import core  # noqa: E402
from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF  # noqa: E402
from oracle.hull import PandaHullChecker  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
margin = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
model = core.load_model("panda_spheres")
lim = np.array(model["joint_limits"])
rng = np.random.default_rng(0)
q = np.round(rng.uniform(lim[:, 0], lim[:, 1], (n, 7)), 2)
base = core.base_matrix((0, 0, 0.01))
scene = [core.plane((0, 0, 1), 0.0), core.box((0.45, 0.1, 0.25), (0.05, 0.2, 0.25)),
         core.box((-0.3, -0.45, 0.5), (0.1, 0.1, 0.1), R=core.rpy_origin((0, 0, 0), (0.3, 0.2, 0.7))[:3, :3]),
         core.sphere((0.3, -0.4, 0.7), 0.12), core.cylinder((-0.4, 0.4, 0.3), (0, 0, 1), 0.3, 0.06)]
hull = PandaHullChecker([base], scene, margin=margin)
q32 = core.pack_f32(core.to_device_f64(q))
print(f"# {n} uniform configurations in the Panda joint limits (rounded to 2 dp), base z = 0.01; "
      f"oracle hulls inflated by {margin * 1000:.1f} mm per body")
print(f"# sphere model: {len(model['spheres'])} spheres, per-link protrusion "
      f"{ {k: round(v['margin_m'] * 1000, 1) for k, v in model['fit_report'].items()} } mm")


def report(name, got, ref):
    fn, fp = int((ref & ~got).sum()), int((~ref & got).sum())
    print(f"{name:52s} hull rate {ref.mean():.4f}  model rate {got.mean():.4f}  false-neg {fn} ({fn / n:.5f})  "
          f"false-pos {fp} ({fp / n:.5f})")


ref_self = hull.self_collision(q)
for tables in (True, False):
    w = core.SpatialWorld([core.RobotSpec(model, base, 0, use_pair_tables=tables)], scene)
    got = w.collide_configs(q32, 0, CHECK_SELF).cpu().numpy().astype(bool)
    report(f"self collision ({'proven-free pairs + 5/7 table' if tables else 'plain sphere model'})", got, ref_self)
w = core.SpatialWorld([core.RobotSpec(model, base, 0)], scene)
ref_obs = hull.obstacle_collision(q)
report("obstacles (plane, 2 boxes, sphere, cylinder)", w.collide_configs(q32, 0, CHECK_OBSTACLES).cpu().numpy().astype(bool), ref_obs)
for i, ob in enumerate(scene):
    wi = core.SpatialWorld([core.RobotSpec(model, base, 0)], [ob])
    hi = PandaHullChecker([base], [ob], margin=margin)
    report(f"  obstacle {i} ({ob['type']})", wi.collide_configs(q32, 0, CHECK_OBSTACLES).cpu().numpy().astype(bool),
           hi.obstacle_collision(q))
report("is_collision_free_sample negated (self + obstacles)",
       w.collide_configs(q32, 0, CHECK_SELF | CHECK_OBSTACLES).cpu().numpy().astype(bool), ref_self | ref_obs)
# robot-robot, sims/cbs_prm.py:27-28 bases
bases = [core.base_matrix((0.7, 0, 0.01), (0, 0, 1, 0)), core.base_matrix((-0.7, 0, 0.01))]
w2 = core.SpatialWorld([core.RobotSpec(model, bases[0], 0), core.RobotSpec(model, bases[1], 7)], [])
h2 = PandaHullChecker(bases, [], margin=margin)
m = min(n, 4000)
qa, qb = q[:m], np.round(rng.uniform(lim[:, 0], lim[:, 1], (m, 7)), 2)
got = w2.collide_robot_pairs(core.pack_f32(core.to_device_f64(qa)), core.pack_f32(core.to_device_f64(qb)), 0, 1)
n = m
report("robot-robot, bases 1.4 m apart (sims/cbs_prm.py:27-28)", got.cpu().numpy().astype(bool),
       h2.robot_robot_collision(qa, qb, 0, 1))
