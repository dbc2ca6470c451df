#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE ITSELF in the build container.

    python tools/gen_golden.py        (needs /root/reference; never runs on the GPU box)

The reference imports pybullet and matplotlib at module top; neither is installed,
and neither is needed by the code on the planar path or by Panda's FK/metric
methods, so two `sys.modules` stubs are installed first (pybullet exposes only the
JOINT_* integers utils/pb_joint_utils.py:15-23 reads at import; matplotlib.patches
provides Rectangle/Circle with get_bbox().bounds / .center / .radius, which is all
planners/prm/planar/env.py uses).  Everything written below is OUTPUT OF REFERENCE
CODE on seeded inputs:

  planar_predicates.npz  Environment predicates + coupled/decoupled is_collision_free_{sample,edge}
  planar_roadmaps.npz    DistancePRM / KNNPRM / KDTree* / coupled / CBSPRM.add_n_samples roadmaps
                         built by the reference from fixed samples
  panda_fk.npz           Panda.positions_from_fk / forward_kinematics / task_space_pose_distance
  panda_heap_knn.npz     DecoupledPRM.connect_nearest_neighbors_degree / _distance run unbound
                         on a fake self (FK metric and L2), incl. the 12-node known answer
  scipy_knn.npz          res/images/{random,halton}_samples.csv and their k=5 / r=0.2 roadmaps
  roadmap_766.npz        res/images/task_space_roadmap2_adjusted_pruned.csv parsed by the
                         reference's own utils/pb_data_utils.load_roadmap
  task_space_sampler.npz DecoupledPRM.sample_task_space_position outputs + the draws it consumed
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("MMMP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def install_stubs():
    pb = types.ModuleType("pybullet")
    for i, n in enumerate(["JOINT_REVOLUTE", "JOINT_PRISMATIC", "JOINT_SPHERICAL", "JOINT_PLANAR", "JOINT_FIXED",
                           "JOINT_POINT2POINT", "JOINT_GEAR"]):
        setattr(pb, n, i)
    pb.GUI, pb.DIRECT = 1, 2
    sys.modules["pybullet"] = pb
    sys.modules["pybullet_data"] = types.ModuleType("pybullet_data")
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    patches = types.ModuleType("matplotlib.patches")

    class _BBox:
        def __init__(self, x0, y0, x1, y1):
            self._p = np.array([[x0, y0], [x1, y1]], dtype=float)

        def get_points(self):
            return self._p

        @property
        def bounds(self):
            (x0, y0), (x1, y1) = self._p
            return (x0, y0, x1 - x0, y1 - y0)

    class Rectangle:
        def __init__(self, xy, width, height, **kw):
            self.xy, self.width, self.height = xy, width, height

        def get_bbox(self):
            x, y = self.xy
            return _BBox(x, y, x + self.width, y + self.height)

    class Circle:
        def __init__(self, xy, radius, **kw):
            self.center, self.radius = np.array(xy, dtype=float), radius

    class Polygon:
        pass

    patches.Rectangle, patches.Circle, patches.Polygon = Rectangle, Circle, Polygon
    mpl.pyplot, mpl.patches = plt, patches
    for name in ("animation", "cm", "colors", "lines"):
        sub = types.ModuleType("matplotlib." + name)
        setattr(mpl, name, sub)
        sys.modules["matplotlib." + name] = sub
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    sys.modules["matplotlib.patches"] = patches
    sns = types.ModuleType("seaborn")
    sys.modules["seaborn"] = sns
    sys.modules["imageio"] = types.ModuleType("imageio")
    return patches


def planar_world(n_arms, with_obstacles, patches):
    from planners.prm.planar.env import Environment
    from robots.planar_arm import PlanarArm
    # experiments/planar/exp_scalability.py:39-48 layout: 2-link arms of length 3
    bases = [(-5.6, -4.0), (5.6, -4.0), (0.0, 4.5), (-5.6, 5.0)][:n_arms]
    arms = [PlanarArm(np.array([3.0, 3.0]), np.array(b)) for b in bases]
    starts = [np.array([1.2, 0.4]), np.array([2.0, -0.4]), np.array([-1.6, 0.3]), np.array([-0.5, 0.5])]
    agents = [{"name": f"a{i}", "start": starts[i], "goal": starts[i] + 0.3, "model": arms[i]} for i in range(n_arms)]
    obstacles = []
    if with_obstacles:
        obstacles = [patches.Rectangle((-1.5, -2.0), 3.0, 1.2), patches.Rectangle((2.1, 3.3), 1.7, 2.9),
                     patches.Circle((-3.0, 1.5), 1.1)]
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        env = Environment((20, 20), agents, obstacles)
    desc = {"bases": np.array(bases), "links": np.array([[3.0, 3.0]] * n_arms),
            "rects": np.array([[r.get_bbox().bounds[0], r.get_bbox().bounds[1],
                                r.get_bbox().bounds[0] + r.get_bbox().bounds[2],
                                r.get_bbox().bounds[1] + r.get_bbox().bounds[3]]
                               for r in obstacles if isinstance(r, patches.Rectangle)]).reshape(-1, 4),
            "circles": np.array([[c.center[0], c.center[1], c.radius] for c in obstacles
                                 if isinstance(c, patches.Circle)]).reshape(-1, 3)}
    return env, desc


def gen_planar(patches):
    import contextlib
    import io
    from planners.prm.planar.multi_arm.coupled.distance_coupled_prm import DistanceCoupledPRM
    from planners.prm.planar.multi_arm.coupled.degree_coupled_prm import KDTreeKNNCoupledPRM
    from planners.prm.planar.multi_arm.coupled.coupled_prm import CoupledPRM
    from planners.prm.planar.multi_arm.decoupled.cbs_prm import CBSPRM
    from planners.prm.planar.single_arm.distance_prm import DistancePRM, KDTreeDistancePRM
    from planners.prm.planar.single_arm.degree_prm import KNNPRM, KDTreeKNNPRM
    from planners.prm.planar.single_arm.prm import PRM
    rng = np.random.default_rng(1)
    out = {}
    # ---- predicates: 2 arms + obstacles, composite 4-D configs -----------------
    env, desc = planar_world(2, True, patches)
    for k, v in desc.items():
        out["w2_" + k] = v
    cp = CoupledPRM(env)
    q = np.round(rng.uniform(-np.pi, np.pi, (3000, 4)), 2)
    with contextlib.redirect_stdout(io.StringIO()):
        out["coupled_q"] = q
        out["coupled_free"] = np.array([cp.is_collision_free_sample(s) for s in q])
        ea, eb = q[:400], q[400:800]
        out["coupled_edge_free"] = np.array([cp.is_collision_free_edge(a, b) for a, b in zip(ea, eb)])
        # near neighbours make mostly-free edges: perturbations
        eb2 = np.round(ea + rng.normal(0, 0.15, ea.shape), 2)
        out["coupled_edge2_b"] = eb2
        out["coupled_edge2_free"] = np.array([cp.is_collision_free_edge(a, b) for a, b in zip(ea, eb2)])
        # decoupled (bounds + obstacles only) and single-arm (bounds + self + obstacles) wrappers
        from planners.prm.planar.multi_arm.decoupled.decoupled_prm import DecoupledPRM
        dp = DecoupledPRM(env)
        q2 = np.round(rng.uniform(-np.pi, np.pi, (2000, 2)), 2)
        out["dec_q"] = q2
        out["dec_free_r0"] = np.array([dp.is_collision_free_sample(s, 0) for s in q2])
        out["dec_free_r1"] = np.array([dp.is_collision_free_sample(s, 1) for s in q2])
        out["dec_edge_free_r1"] = np.array([dp.is_collision_free_edge(a, b, 1) for a, b in zip(q2[:300], q2[300:600])])
        # robot-robot only
        out["rr_hit"] = np.array([bool(cp.robot_robot_collision(s, env.robot_models[0], env.robot_models[1]))
                                  for s in q[:1500]])
    # ---- single 3-link arm (self collision matters) ------------------------------
    from planners.prm.planar.env import Environment
    from robots.planar_arm import PlanarArm
    arm3 = PlanarArm(np.array([2.5, 2.0, 2.2]), np.array([0.5, -1.0]))
    obs3 = [patches.Rectangle((3.0, 1.0), 2.0, 2.0), patches.Circle((-2.5, 2.5), 1.0)]
    with contextlib.redirect_stdout(io.StringIO()):
        env3 = Environment((20, 20), [{"name": "a", "start": np.array([0.1, 0.2, 0.3]),
                                       "goal": np.array([0.3, 0.2, 0.1]), "model": arm3}], obs3)
        p3 = PRM(env3, 0.05)
        q3 = np.round(rng.uniform(-np.pi, np.pi, (3000, 3)), 2)
        out["single_q"] = q3
        out["single_free"] = np.array([p3.is_collision_free_sample(s) for s in q3])
        out["single_fk"] = np.array([arm3.forward_kinematics(s) for s in q3[:200]])
        out["single_edge_free"] = np.array([p3.is_collision_free_edge(a, b) for a, b in zip(q3[:300], q3[300:600])])
        out["single_specific_distance"] = np.array([p3.specific_distance(a, b) for a, b in zip(q3[:300], q3[300:600])])
    np.savez_compressed(os.path.join(OUT, "planar_predicates.npz"), **out)

    # ---- roadmaps built by the reference from FIXED samples --------------------
    rm = {}

    def feed(planner_cls_init, samples):
        """Build a planner whose generate_free_sample pops from `samples`."""
        it = iter(samples)
        planner = planner_cls_init()
        planner.generate_free_sample = lambda *a: next(it)
        return planner

    def adj(roadmap, n):
        return np.array([len(roadmap[i]) for i in range(n)]), np.concatenate([np.array(roadmap[i], dtype=np.int64)
                                                                              for i in range(n)] + [np.zeros(0, np.int64)])

    with contextlib.redirect_stdout(io.StringIO()):
        # free samples for the single arm
        free3 = q3[out["single_free"]][:150]
        rm["single_samples"] = free3
        for name, cls, kw in (("distance", DistancePRM, dict(max_edge_len=6.0)),
                              ("knn", KNNPRM, dict(max_edge_len=6.0, n_knn=4))):
            pl = cls.__new__(cls)
            PRM.__init__(pl, env3, 0.05)
            for k_, v_ in kw.items():
                setattr(pl, k_, v_)
            it = iter(free3)
            pl.generate_free_sample = lambda it=it: next(it)
            pl.add_n_samples(len(free3))
            rm[name + "_deg"], rm[name + "_adj"] = adj(pl.roadmap, len(free3))
        for name, cls, kw in (("kd_distance", KDTreeDistancePRM, dict(max_edge_len=1.2)),
                              ("kd_knn", KDTreeKNNPRM, dict(max_edge_len=1.2, n_knn=4))):
            pl = cls.__new__(cls)
            PRM.__init__(pl, env3, 0.05)
            for k_, v_ in kw.items():
                setattr(pl, k_, v_)
            it = iter(free3)
            pl.generate_free_sample = lambda it=it: next(it)
            pl.add_n_samples(len(free3))
            pl.update_roadmap()
            rm[name + "_deg"], rm[name + "_adj"] = adj(pl.roadmap, len(free3))
        # coupled 2 arms
        free4 = q[out["coupled_free"]][:200]
        rm["coupled_samples"] = free4
        pl = DistanceCoupledPRM.__new__(DistanceCoupledPRM)
        CoupledPRM.__init__(pl, env)
        pl.max_edge_len = 1.6
        it = iter(free4)
        pl.generate_free_sample = lambda it=it: next(it)
        pl.add_n_samples(len(free4))
        rm["coupled_distance_deg"], rm["coupled_distance_adj"] = adj(pl.roadmap, len(free4))
        pl = KDTreeKNNCoupledPRM.__new__(KDTreeKNNCoupledPRM)
        CoupledPRM.__init__(pl, env)
        pl.max_edge_len, pl.n_knn = 1.6, 5
        it = iter(free4)
        pl.generate_free_sample = lambda it=it: next(it)
        pl.add_n_samples(len(free4))
        pl.update_roadmap()
        rm["coupled_kdknn_deg"], rm["coupled_kdknn_adj"] = adj(pl.roadmap, len(free4))
        # decoupled CBS learning phase, arm 1 (obstacles near it)
        free2 = q2[out["dec_free_r1"]][:250]
        rm["cbs_samples_r1"] = free2
        pl = CBSPRM.__new__(CBSPRM)
        from planners.prm.planar.multi_arm.decoupled.decoupled_prm import DecoupledPRM
        DecoupledPRM.__init__(pl, env)
        pl.max_edge_len = 0.9
        it = iter(free2)
        pl.generate_free_sample = lambda rid, it=it: next(it)
        pl.add_n_samples(len(free2), 1)
        rm["cbs_deg_r1"], rm["cbs_adj_r1"] = adj(pl.roadmaps[1], len(free2))
    np.savez_compressed(os.path.join(OUT, "planar_roadmaps.npz"), **rm)
    print("planar:", {k: (v.shape, float(np.mean(v)) if v.dtype == bool else None) for k, v in out.items()
                      if k.endswith("free") or k.endswith("hit")})


def make_panda(base_position=(0, 0, 0), base_orientation=(0, 0, 0, 1)):
    from robots.panda import Panda
    p = Panda.__new__(Panda)
    p.base_position, p.base_orientation = base_position, base_orientation
    p.dh_params = [
        {"a": 0, "d": 0.333, "alpha": 0, "theta": 0}, {"a": 0, "d": 0, "alpha": -np.pi / 2, "theta": 0},
        {"a": 0, "d": 0.316, "alpha": np.pi / 2, "theta": 0}, {"a": 0.0825, "d": 0, "alpha": np.pi / 2, "theta": 0},
        {"a": -0.0825, "d": 0.384, "alpha": -np.pi / 2, "theta": 0}, {"a": 0, "d": 0, "alpha": np.pi / 2, "theta": 0},
        {"a": 0.088, "d": 0, "alpha": np.pi / 2, "theta": 0}, {"a": 0, "d": 0.107, "alpha": 0, "theta": 0}]
    p.arm_dimension = 7
    return p


LIMITS = np.array([(-2.8973, 2.8973), (-1.7628, 1.7628), (-2.8973, 2.8973), (-3.0718, -0.0698),
                   (-2.8973, 2.8973), (-0.0175, 3.7525), (-2.8973, 2.8973)])


def gen_panda():
    import heapq
    from planners.prm.pybullet.decoupled.decoupled_prm import DecoupledPRM
    from planners.prm.pybullet.utils import Node
    rng = np.random.default_rng(2)
    out = {}
    for tag, pos, orn in (("id", (0, 0, 0), (0, 0, 0, 1)), ("b1", (0.7, 0.0, 0.01), (0, 0, 1, 0)),
                          ("b2", (-0.3, 0.4, 0.02), (0, 0, 0.38268343, 0.92387953))):
        p = make_panda(pos, orn)
        q = np.round(rng.uniform(LIMITS[:, 0], LIMITS[:, 1], (400, 7)), 2)
        out[f"fk_{tag}_q"] = q
        out[f"fk_{tag}_pos"] = np.array([np.array(p.positions_from_fk(s)) for s in q])
        out[f"fk_{tag}_T"] = np.array([p.forward_kinematics(s) for s in q])
        out[f"fk_{tag}_base"] = np.array(list(pos) + list(orn))
    p = make_panda()
    q = out["fk_id_q"]
    out["pose_dist"] = np.array([p.task_space_pose_distance(a, b) for a, b in zip(q[:200], q[200:])])
    q_ref = [0, -np.pi / 4, 0, -3 * np.pi / 4, 0, np.pi / 2, np.pi / 4]
    out["known_dist"] = np.array([p.task_space_pose_distance(q_ref, [.1, -.5, .2, -2, .3, 1.5, .7])])
    np.savez_compressed(os.path.join(OUT, "panda_fk.npz"), **out)

    # ---- production connect methods run unbound on a fake self --------------------
    class Fake:
        pass

    def run(nodes_q, metric, k1, k2, maxdist, edge_free):
        f = Fake()
        f.r_ids = ["r"]
        f.node_lists = [[Node(i, s) for i, s in enumerate(nodes_q)]]
        f.k1, f.k2, f.maxdist = k1, k2, maxdist
        f.heapsort = lambda it: DecoupledPRM.heapsort(f, it)
        if metric == "fk":
            f.distance = lambda r, a, b: p.task_space_pose_distance(a, b)
        else:
            f.distance = lambda r, a, b: np.linalg.norm(np.array(a) - np.array(b))
        f.is_collision_free_edge = edge_free
        res = {}
        for mode in ("degree", "distance"):
            f.edge_dicts = [{n.id: [] for n in f.node_lists[0]}]
            getattr(DecoupledPRM, "connect_nearest_neighbors_" + mode)(f, "r")
            ed = f.edge_dicts[0]
            res[mode + "_deg"] = np.array([len(ed[i]) for i in range(len(nodes_q))])
            res[mode + "_adj"] = np.concatenate([np.array(ed[i], dtype=np.int64) for i in range(len(nodes_q))] +
                                                [np.zeros(0, np.int64)])
        return res

    kk = {}
    n12 = [[1, 1], [1, 2], [2, 1], [10, 10], [11, 10], [10, 11], [20, 20], [21, 20], [20, 21], [30, 30], [30, 31],
           [31, 30]]
    for k_, v_ in run(n12, "l2", 5, 3, 1.5, lambda a, b, r: True).items():
        kk["n12_" + k_] = v_
    qs = np.round(rng.uniform(LIMITS[:, 0], LIMITS[:, 1], (300, 7)), 2)
    kk["samples"] = qs
    # pseudo-random but deterministic edge verdicts (a pure function of the endpoints)
    def ef(a, b, r):
        return (int(round((sum(a) + sum(b)) * 100)) % 7) != 0
    kk["edge_rule"] = np.array([7])
    for metric in ("fk", "l2"):
        for k_, v_ in run(qs, metric, 10, 6, 2.0 if metric == "fk" else 1.6, ef).items():
            kk[f"{metric}_{k_}"] = v_
        # candidate lists themselves (heap loop only): k1 nearest by (dist, id)
        f_dist = (lambda a, b: p.task_space_pose_distance(a, b)) if metric == "fk" else \
            (lambda a, b: np.linalg.norm(np.array(a) - np.array(b)))
        nodes = [Node(i, s) for i, s in enumerate(qs)]
        idx = np.zeros((len(qs), 10), np.int64)
        dst = np.zeros((len(qs), 10))
        for i, nd in enumerate(nodes):
            h = []
            for other in nodes:
                if nd == other:
                    continue
                heapq.heappush(h, (-f_dist(nd.q, other.q), other))
                if len(h) > 10:
                    heapq.heappop(h)
            c = sorted([(-d, o) for d, o in h])
            idx[i] = [o.id for _, o in c]
            dst[i] = [d for d, _ in c]
        kk[f"{metric}_knn_idx"], kk[f"{metric}_knn_dist"] = idx, dst
    np.savez_compressed(os.path.join(OUT, "panda_heap_knn.npz"), **kk)
    print("panda: known_dist", out["known_dist"], "n12 degree adj", kk["n12_degree_adj"][:12])


def gen_fixtures():
    from utils.pb_data_utils import load_roadmap
    d = os.path.join(REF, "res/images")
    out = {}
    for name in ("random", "halton"):
        out[name + "_samples"] = np.loadtxt(os.path.join(d, name + "_samples.csv"), delimiter=",")
        out[name + "_degree"] = np.loadtxt(os.path.join(d, name + "_degree_roadmap.csv"), delimiter=",", dtype=np.int64)
        out[name + "_distance"] = np.loadtxt(os.path.join(d, name + "_distance_roadmap.csv"), delimiter=",",
                                             dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "scipy_knn.npz"), **out)
    rm = load_roadmap(os.path.join(d, "task_space_roadmap2_adjusted_pruned.csv"))
    keys = list(rm.keys())
    q = np.array(keys, dtype=np.float64)
    qr = np.round(q, 2)
    ids = {tuple(k): i for i, k in enumerate(qr)}
    deg = np.array([len(rm[k]) for k in keys])
    adj = np.array([ids[tuple(np.round(np.array(nb), 2))] for k in keys for nb in rm[k]], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "roadmap_766.npz"), q=q, deg=deg, adj=adj)
    print("fixtures: 766-roadmap nodes", len(keys), "directed edges", len(adj))


def gen_task_space_sampler():
    """DecoupledPRM.sample_task_space_position (decoupled_prm.py:346-369), called unbound (it
    does not touch self) under a seeded numpy global stream; the draws it consumed are
    recovered by replaying the same stream (uniform, chisquare(1), uniform per sample)."""
    from planners.prm.pybullet.decoupled.decoupled_prm import DecoupledPRM
    out = {}
    cases = [((0.31, -0.2, 0.45), (0.55, 0.35, 0.3)), ((0.2, 0.0, 0.5), (0.7, 0.0, 0.5)),   # second: W = +x
             ((-0.4, 0.3, 0.2), (0.1, -0.5, 0.8))]
    for c, (start, goal) in enumerate(cases):
        np.random.seed(100 + c)
        pos = np.array([DecoupledPRM.sample_task_space_position(None, start, goal) for _ in range(200)])
        np.random.seed(100 + c)
        draws = np.array([[np.random.uniform(0, 1), np.random.chisquare(1), np.random.uniform(0, 2 * np.pi)]
                          for _ in range(200)])
        out[f"start{c}"], out[f"goal{c}"], out[f"pos{c}"], out[f"draws{c}"] = np.array(start), np.array(goal), pos, draws
    np.savez_compressed(os.path.join(OUT, "task_space_sampler.npz"), **out)
    print("task-space sampler: first position", out["pos0"][0])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    patches = install_stubs()
    sys.path.insert(0, REF)
    os.chdir(REF)
    only = sys.argv[1:]
    if not only or "planar" in only:
        gen_planar(patches)
    if not only or "panda" in only:
        gen_panda()
    if not only or "fixtures" in only:
        gen_fixtures()
    if not only or "task_space" in only:
        gen_task_space_sampler()
