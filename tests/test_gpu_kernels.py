"""GPU parity tests: every kernel, called through the C ABI, against the oracle
(oracle/, CPU float64) on identical seeded inputs and against the golden vectors the
reference itself produced (tests/golden/, tools/gen_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIMITS = np.array([(-2.8973, 2.8973), (-1.7628, 1.7628), (-2.8973, 2.8973), (-3.0718, -0.0698),
                   (-2.8973, 2.8973), (-0.0175, 3.7525), (-2.8973, 2.8973)])


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def core(cuda):
    from mmmp_b200 import core
    return core


@pytest.fixture(scope="module")
def panda_model(core):
    return core.load_model("panda_spheres")


# --------------------------------------------------------------------- (a) sampling
def test_sampler_bit_exact_vs_oracle(core):
    from oracle import sampling
    for D, seed, first in ((7, 1234, 0), (14, 99, 10**6), (2, 5, 2**33 + 17), (28, 2**40 + 3, 7)):
        lo = np.linspace(-3, -1, D)
        hi = np.linspace(0.5, 3, D)
        q64, q32 = core.sample_uniform(lo, hi, 5000, seed, first)
        ref = sampling.sample_uniform(lo, hi, 5000, seed, first)
        assert (_np(q64) == ref).all()
        assert (_np(q32)[:, :D] == ref.astype(np.float32)).all()
        assert (_np(q32)[:, D:] == 0).all()
    q64, _ = core.sample_uniform([0.0], [1.0], 200000, 3, 0, round_dp=-1)
    u = _np(q64)[:, 0]
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3


def test_compaction_and_gather(core, cuda):
    rng = np.random.default_rng(0)
    for n in (1, 7, 2048, 2049, 100003):
        f = (rng.random(n) < 0.37).astype(np.uint8)
        idx = core.compact_free(torch.from_numpy(f).to(cuda))
        assert (_np(idx) == np.nonzero(f == 0)[0]).all()
    t = torch.arange(40000, dtype=torch.float64, device=cuda).reshape(-1, 4)
    idx = torch.tensor([5, 0, 9999, 17], dtype=torch.int32, device=cuda)
    assert (core.gather_rows(t, idx) == t[idx.long()]).all()
    c = torch.from_numpy(rng.integers(0, 9, 70001).astype(np.int32)).to(cuda)
    off = core.exclusive_scan(c)
    assert (_np(off) == np.concatenate([[0], np.cumsum(_np(c).astype(np.int64))])).all()


# --------------------------------------------------------------------- (b) FK
def test_panda_fk_vs_reference_golden(core, golden, panda_model, cuda):
    g = golden("panda_fk")
    for tag in ("id", "b1", "b2"):
        b = g[f"fk_{tag}_base"]
        w = core.SpatialWorld([core.RobotSpec(panda_model, core.base_matrix(b[:3], b[3:]))])
        q64 = core.to_device_f64(g[f"fk_{tag}_q"])
        q32 = core.pack_f32(q64)
        pos, T = w.fk(q32, 0, poses=True)
        assert np.abs(_np(pos) - g[f"fk_{tag}_pos"]).max() < 1e-5          # north_star tolerance
        Tref = g[f"fk_{tag}_T"][:, :3, :].reshape(-1, 12)
        assert np.abs(_np(T)[:, 7] - Tref).max() < 1e-5
        pos64 = w.fk64(q64, 0)
        assert np.abs(_np(pos64) - g[f"fk_{tag}_pos"]).max() < 1e-13
        w.close()


def test_kuka_chain_vs_oracle(core, cuda):
    from oracle import chain
    from oracle.spheres import model_origins
    m = core.load_model("kuka_iiwa14_spheres")
    rng = np.random.default_rng(3)
    lim = np.array(m["joint_limits"])
    q = rng.uniform(lim[:, 0], lim[:, 1], (500, 7))
    base = core.base_matrix((0.2, -0.1, 0.3), (0, 0, 0.6, 0.8))
    w = core.SpatialWorld([core.RobotSpec(m, base)])
    pos = w.fk(core.pack_f32(core.to_device_f64(q)), 0)
    ref = chain.chain_frames(model_origins(m), q, base)[:, :, :3, 3]
    assert np.abs(_np(pos) - ref).max() < 1e-5


def test_planar_fk_vs_reference_golden(core, golden, cuda):
    g = golden("planar_predicates")
    w = core.PlanarWorld([{"links": [2.5, 2.0, 2.2], "base": (0.5, -1.0)}], (-10, 10, -10, 10))
    pos = w.fk(core.to_device_f64(g["single_q"][:200]), 0)
    assert np.abs(_np(pos) - g["single_fk"]).max() < 1e-14
    # reference tests/robots/test_planar_arm.py:19-26
    w2 = core.PlanarWorld([{"links": [1.0, 1.0], "base": (0.0, 0.0)}], (-10, 10, -10, 10))
    pos = w2.fk(core.to_device_f64([[np.pi / 2, np.pi / 2]]), 0)
    np.testing.assert_almost_equal(_np(pos)[0], np.array([[0, 1], [-1, 1]]), decimal=5)


# --------------------------------------------------------------------- (c) planar collision
def _planar_world(core, g):
    rects = [{"type": "rect", "x0": r[0], "y0": r[1], "x1": r[2], "y1": r[3]} for r in g["w2_rects"]]
    circ = [{"type": "circle", "cx": c[0], "cy": c[1], "r": c[2]} for c in g["w2_circles"]]
    arms = [{"links": list(g["w2_links"][i]), "base": tuple(g["w2_bases"][i])} for i in range(2)]
    return core.PlanarWorld(arms, (-10, 10, -10, 10), rects + circ)


def test_planar_collision_vs_reference_golden(core, golden, cuda):
    from mmmp_b200._lib import CHECK_BOUNDS, CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF
    g = golden("planar_predicates")
    w = _planar_world(core, g)
    ALL = CHECK_BOUNDS | CHECK_OBSTACLES | CHECK_ROBOTS | CHECK_SELF
    q = core.to_device_f64(g["coupled_q"])
    hit = _np(w.collide_configs(q, -1, ALL)).astype(bool)
    assert (hit == ~g["coupled_free"]).all()                       # exact: 3000 composite configs
    rr = _np(w.collide_configs(q[:1500].contiguous(), -1, CHECK_ROBOTS)).astype(bool)
    assert (rr == g["rr_hit"]).all()
    q2 = core.to_device_f64(g["dec_q"])
    for arm, key in ((0, "dec_free_r0"), (1, "dec_free_r1")):
        hit = _np(w.collide_configs(q2, arm, CHECK_BOUNDS | CHECK_OBSTACLES)).astype(bool)
        assert (hit == ~g[key]).all()
    # edges: is_collision_free_edge with LOCAL_STEP 0.05 -> 20 steps
    a, b = q[:400].contiguous(), q[400:800].contiguous()
    eh = _np(w.collide_segments(a, b, 20, 0.05, -1, ALL)).astype(bool)
    assert (eh == ~g["coupled_edge_free"]).all()
    b2 = core.to_device_f64(g["coupled_edge2_b"])
    eh = _np(w.collide_segments(a, b2, 20, 0.05, -1, ALL)).astype(bool)
    assert (eh == ~g["coupled_edge2_free"]).all()
    e = torch.stack([torch.arange(300), torch.arange(300, 600)], 1).to(torch.int32).to(cuda).contiguous()
    eh = _np(w.collide_edges(q2, e, 20, 0.05, 1, CHECK_BOUNDS | CHECK_OBSTACLES)).astype(bool)
    assert (eh == ~g["dec_edge_free_r1"]).all()
    # single 3-link arm: bounds + self + obstacles
    w3 = core.PlanarWorld([{"links": [2.5, 2.0, 2.2], "base": (0.5, -1.0)}], (-10, 10, -10, 10),
                          [{"type": "rect", "x0": 3.0, "y0": 1.0, "x1": 5.0, "y1": 3.0},
                           {"type": "circle", "cx": -2.5, "cy": 2.5, "r": 1.0}])
    q3 = core.to_device_f64(g["single_q"])
    F3 = CHECK_BOUNDS | CHECK_OBSTACLES | CHECK_SELF
    assert (_np(w3.collide_configs(q3, 0, F3)).astype(bool) == ~g["single_free"]).all()
    eh = _np(w3.collide_segments(q3[:300].contiguous(), q3[300:600].contiguous(), 20, 0.05, 0, F3)).astype(bool)
    assert (eh == ~g["single_edge_free"]).all()


# --------------------------------------------------------------------- (c) spatial collision
def _scene(core):
    return [core.plane((0, 0, 1), 0.0),
            core.box((0.45, 0.1, 0.25), (0.05, 0.2, 0.25)),
            core.box((-0.3, -0.45, 0.5), (0.1, 0.1, 0.1), R=core.rpy_origin((0, 0, 0), (0.3, 0.2, 0.7))[:3, :3]),
            core.sphere((0.3, -0.4, 0.7), 0.12),
            core.cylinder((-0.4, 0.4, 0.3), (0, 0, 1), 0.3, 0.06)]


def test_panda_config_collision_vs_sphere_oracle(core, panda_model, cuda):
    from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_SELF
    from oracle.spheres import SphereChecker
    rng = np.random.default_rng(11)
    base = core.base_matrix((0, 0, 0.01))
    obs = _scene(core)
    w = core.SpatialWorld([core.RobotSpec(panda_model, base)], obs)
    q = np.round(rng.uniform(LIMITS[:, 0], LIMITS[:, 1], (6000, 7)), 2)
    q32 = core.pack_f32(core.to_device_f64(q))
    ck = SphereChecker([panda_model], [base], obs)
    for flags, kw in ((CHECK_SELF, dict(self_=True, obstacles=False)), (CHECK_OBSTACLES, dict(self_=False, obstacles=True)),
                      (CHECK_SELF | CHECK_OBSTACLES, dict(self_=True, obstacles=True))):
        got = _np(w.collide_configs(q32, 0, flags)).astype(bool)
        ref = ck.config_collision(q, 0, **kw)
        mism = np.nonzero(got != ref)[0]
        # fp32 vs float64 can only disagree on a sphere pair within ~1e-6 m of touching
        assert len(mism) <= 3, (flags, len(mism))
        assert 0.02 < ref.mean() < 0.98
    # no pair table -> pure sphere model for link5/link7 as well
    w2 = core.SpatialWorld([core.RobotSpec(panda_model, base, use_pair_tables=False)], obs)
    ck2 = SphereChecker([panda_model], [base], obs, use_tables=False)
    got = _np(w2.collide_configs(q32, 0, CHECK_SELF)).astype(bool)
    assert (got != ck2.config_collision(q, 0, self_=True, obstacles=False)).sum() <= 3


def test_panda_edges_vs_sphere_oracle(core, panda_model, cuda):
    from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_SELF
    from oracle.spheres import SphereChecker
    rng = np.random.default_rng(12)
    base = core.base_matrix((0, 0, 0.01))
    obs = _scene(core)
    w = core.SpatialWorld([core.RobotSpec(panda_model, base)], obs)
    ck = SphereChecker([panda_model], [base], obs)
    qa = np.round(rng.uniform(LIMITS[:, 0], LIMITS[:, 1], (700, 7)), 2)
    qb = np.clip(np.round(qa + rng.normal(0, 0.25, qa.shape), 2), LIMITS[:, 0], LIMITS[:, 1])
    F = CHECK_SELF | CHECK_OBSTACLES
    for n_steps, step in ((50, 0.02), (20, 0.05), (7, 0.15)):
        ref = ck.edge_collision(qa, qb, n_steps, step, 0)
        a32, b32 = core.pack_f32(core.to_device_f64(qa)), core.pack_f32(core.to_device_f64(qb))
        got = _np(w.collide_segments(a32, b32, n_steps, step, 0, F)).astype(bool)
        assert (got != ref).sum() <= 3
        allq = core.pack_f32(core.to_device_f64(np.concatenate([qa, qb])))
        e = torch.stack([torch.arange(700), torch.arange(700, 1400)], 1).to(torch.int32).to(cuda).contiguous()
        got2 = _np(w.collide_edges(allq, e, n_steps, step, 0, F)).astype(bool)
        assert (got2 == got).all()
        assert 0.05 < ref.mean() < 0.95
    # empty neighbour slots (-1) are reported blocked
    e = torch.tensor([[0, -1], [-1, 3]], dtype=torch.int32, device=cuda)
    assert _np(w.collide_edges(allq, e, 50, 0.02, 0, F)).tolist() == [1, 1]


def test_edge_step_skipping_keeps_every_verdict(core, panda_model, cuda):
    """The default edge kernel skips interpolation steps a clearance bound proves free; its
    verdicts must equal the exhaustive kernel's (every step evaluated) bit for bit."""
    from mmmp_b200._lib import CHECK_EXHAUSTIVE, CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF
    rng = np.random.default_rng(21)
    base = core.base_matrix((0, 0, 0.01))
    n = 30000
    qa = np.round(rng.uniform(LIMITS[:, 0], LIMITS[:, 1], (n, 7)), 2)
    sigma = rng.choice([0.02, 0.1, 0.3, 1.0], size=(n, 1))     # short roadmap edges up to long ones
    qb = np.clip(np.round(qa + rng.normal(0, 1.0, qa.shape) * sigma, 2), LIMITS[:, 0], LIMITS[:, 1])
    qb[:50] = qa[:50]                                            # degenerate edges
    a32, b32 = core.pack_f32(core.to_device_f64(qa)), core.pack_f32(core.to_device_f64(qb))
    for obs in (_scene(core), [core.plane((0, 0, 1), 0.0)], []):
        w = core.SpatialWorld([core.RobotSpec(panda_model, base)], obs)
        for flags in (CHECK_SELF, CHECK_OBSTACLES, CHECK_SELF | CHECK_OBSTACLES):
            for n_steps, step in ((50, 0.02), (64, 1.0 / 64), (20, 0.05), (5, 0.2)):
                fast = _np(w.collide_segments(a32, b32, n_steps, step, 0, flags))
                full = _np(w.collide_segments(a32, b32, n_steps, step, 0, flags | CHECK_EXHAUSTIVE))
                assert (fast == full).all(), (len(obs), flags, n_steps, int((fast != full).sum()))
        assert 0.02 < full.mean() < 0.98
    # composite rows of two robots, robot-robot pairs included
    bases = [core.base_matrix((-0.35, 0, 0.01)), core.base_matrix((0.35, 0, 0.01), (0, 0, 1, 0))]
    w2 = core.SpatialWorld([core.RobotSpec(panda_model, bases[0], 0), core.RobotSpec(panda_model, bases[1], 7)],
                           [core.plane((0, 0, 1), 0.0)])
    m = 6000
    ca = np.concatenate([qa[:m], qa[m:2 * m]], 1)
    cb = np.concatenate([qb[:m], qb[m:2 * m]], 1)
    ca32, cb32 = core.pack_f32(core.to_device_f64(ca)), core.pack_f32(core.to_device_f64(cb))
    for flags in (CHECK_ROBOTS, CHECK_SELF | CHECK_OBSTACLES | CHECK_ROBOTS):
        fast = _np(w2.collide_segments(ca32, cb32, 50, 0.02, -1, flags))
        full = _np(w2.collide_segments(ca32, cb32, 50, 0.02, -1, flags | CHECK_EXHAUSTIVE))
        assert (fast == full).all(), (flags, int((fast != full).sum()))
        assert 0.02 < full.mean() < 0.98


def test_capsule_cull_keeps_every_verdict(core, panda_model, cuda):
    """Between a link pair's bounding spheres and its sphere sets the kernels test the links'
    bounding CAPSULES (fitted at world creation).  The capsules contain every sphere of their
    link, so the cull can only skip descents that would not have hit: verdicts with and without
    the stage (MMMP_CHECK_NO_CAPSULES) are identical -- 2 M single-arm configurations, 400 k
    two-arm composite rows (robot-robot pairs), KUKA, and 200 k edges through the step-skipping
    and the exhaustive kernels (whose clearances the capsules change, but not their OR)."""
    from mmmp_b200._lib import CHECK_EXHAUSTIVE, CHECK_NO_CAPSULES, CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF
    NC = CHECK_NO_CAPSULES
    lim = np.array(panda_model["joint_limits"])
    w = core.SpatialWorld([core.RobotSpec(panda_model, core.base_matrix((0, 0, 0.01)))], _scene(core))
    q64, q32 = core.sample_uniform(lim[:, 0], lim[:, 1], 2_000_000, seed=99)
    for flags in (CHECK_SELF, CHECK_SELF | CHECK_OBSTACLES):
        a, b = w.collide_configs(q32, 0, flags), w.collide_configs(q32, 0, flags | NC)
        assert torch.equal(a, b), int((a != b).sum())
        assert 0.05 < float(a.float().mean()) < 0.6
    # the plain sphere model, without the offline pair tables / proven pairs: every self pair goes
    # through the capsule stage
    wp = core.SpatialWorld([core.RobotSpec(panda_model, core.base_matrix((0, 0, 0.01)), use_pair_tables=False)], [])
    a, b = wp.collide_configs(q32, 0, CHECK_SELF), wp.collide_configs(q32, 0, CHECK_SELF | NC)
    assert torch.equal(a, b), int((a != b).sum())
    # edges: roadmap-like short edges and long ones
    rng = np.random.default_rng(5)
    n = 200_000
    qa = _np(q64[:n])
    sigma = rng.choice([0.02, 0.1, 0.3, 1.0], size=(n, 1))
    qb = np.clip(np.round(qa + rng.normal(0, 1.0, qa.shape) * sigma, 2), lim[:, 0], lim[:, 1])
    a32, b32 = q32[:n].contiguous(), core.pack_f32(core.to_device_f64(qb))
    F = CHECK_SELF | CHECK_OBSTACLES
    ref = w.collide_segments(a32, b32, 50, 0.02, 0, F | NC | CHECK_EXHAUSTIVE)
    for flags in (F, F | CHECK_EXHAUSTIVE, F | NC):
        got = w.collide_segments(a32, b32, 50, 0.02, 0, flags)
        assert torch.equal(got, ref), (flags, int((got != ref).sum()))
    assert 0.05 < float(ref.float().mean()) < 0.95
    # two arms close together: robot-robot link pairs
    bases = [core.base_matrix((-0.35, 0, 0.01)), core.base_matrix((0.35, 0, 0.01), (0, 0, 1, 0))]
    w2 = core.SpatialWorld([core.RobotSpec(panda_model, bases[0], 0), core.RobotSpec(panda_model, bases[1], 7)],
                           [core.plane((0, 0, 1), 0.0)])
    c64, c32 = core.sample_uniform(np.tile(lim[:, 0], 2), np.tile(lim[:, 1], 2), 400_000, seed=3)
    for flags in (CHECK_ROBOTS, CHECK_SELF | CHECK_OBSTACLES | CHECK_ROBOTS):
        a, b = w2.collide_configs(c32, -1, flags), w2.collide_configs(c32, -1, flags | NC)
        assert torch.equal(a, b), int((a != b).sum())
        assert 0.05 < float(a.float().mean()) < 0.95
    m = 40_000
    ea, eb = c32[:m].contiguous(), c32[m:2 * m].contiguous()
    FR = CHECK_SELF | CHECK_OBSTACLES | CHECK_ROBOTS
    ref = w2.collide_segments(ea, eb, 20, 0.05, -1, FR | NC | CHECK_EXHAUSTIVE)
    assert torch.equal(w2.collide_segments(ea, eb, 20, 0.05, -1, FR), ref)
    # KUKA iiwa14 (the reference's own sphere URDF)
    km = core.load_model("kuka_iiwa14_spheres")
    klim = np.array(km["joint_limits"])
    wk = core.SpatialWorld([core.RobotSpec(km, core.base_matrix((0, 0, 0.0)))],
                           [core.plane((0, 0, 1), -0.2), core.sphere((0.4, 0.1, 0.6), 0.2)])
    k64, k32 = core.sample_uniform(klim[:, 0], klim[:, 1], 500_000, seed=4)
    a, b = wk.collide_configs(k32, 0, F), wk.collide_configs(k32, 0, F | NC)
    assert torch.equal(a, b)       # (the URDF's own spheres of links two apart always overlap: all ones)
    a = wk.collide_configs(k32, 0, CHECK_OBSTACLES)
    assert 0.02 < float(a.float().mean()) < 0.98


def test_two_pandas_composite_and_pairs_vs_sphere_oracle(core, panda_model, cuda):
    from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF
    from oracle.spheres import SphereChecker
    rng = np.random.default_rng(13)
    # sims/cbs_prm.py:27-28 layout, closer so that robot-robot collisions are frequent
    bases = [core.base_matrix((-0.35, 0, 0.01)), core.base_matrix((0.35, 0, 0.01), (0, 0, 1, 0))]
    obs = [core.plane((0, 0, 1), 0.0)]
    w = core.SpatialWorld([core.RobotSpec(panda_model, bases[0], 0), core.RobotSpec(panda_model, bases[1], 7)], obs)
    assert w.dof == 14
    ck = SphereChecker([panda_model, panda_model], bases, obs)
    q = np.round(rng.uniform(np.tile(LIMITS[:, 0], 2), np.tile(LIMITS[:, 1], 2), (1500, 14)), 2)
    q32 = core.pack_f32(core.to_device_f64(q))
    rr_ref = ck.robot_robot_collision(q[:, :7], q[:, 7:], 0, 1)
    got = _np(w.collide_configs(q32, -1, CHECK_ROBOTS)).astype(bool)
    assert (got != rr_ref).sum() <= 2 and 0.05 < rr_ref.mean() < 0.95
    qa, qb = core.pack_f32(core.to_device_f64(q[:, :7])), core.pack_f32(core.to_device_f64(q[:, 7:]))
    got = _np(w.collide_robot_pairs(qa, qb, 0, 1)).astype(bool)
    assert (got != rr_ref).sum() <= 2
    full = ck.composite_collision(q)
    got = _np(w.collide_configs(q32, -1, CHECK_SELF | CHECK_OBSTACLES | CHECK_ROBOTS)).astype(bool)
    assert (got != full).sum() <= 3


def test_sphere_model_is_conservative_vs_hull_oracle(core, panda_model, cuda):
    """Disagreement with the convex-hull oracle (the restatement of pybullet's verdict, parity
    unpinned): the sphere cover may only ADD collisions (fit margin), never miss one."""
    from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_SELF
    from oracle.hull import PandaHullChecker
    rng = np.random.default_rng(14)
    base = core.base_matrix((0, 0, 0.01))
    obs = _scene(core)
    w = core.SpatialWorld([core.RobotSpec(panda_model, base)], obs)
    q = np.round(rng.uniform(LIMITS[:, 0], LIMITS[:, 1], (4000, 7)), 2)
    q32 = core.pack_f32(core.to_device_f64(q))
    hull = PandaHullChecker([base], obs)
    for flags, ref in ((CHECK_SELF, hull.self_collision(q)), (CHECK_OBSTACLES, hull.obstacle_collision(q))):
        got = _np(w.collide_configs(q32, 0, flags)).astype(bool)
        false_neg = (ref & ~got).sum()
        false_pos = (~ref & got).mean()
        assert false_neg == 0, f"sphere model missed {false_neg} hull collisions (flags={flags})"
        # measured 3.6 % for the whole accept test on 20 000 configurations (profiles/r01_disagreement.txt:
        # self 2.5 %, obstacles 2.0 %); 4000 samples: 3 sigma of sampling noise on top
        assert false_pos < 0.05


# --------------------------------------------------------------------- (d) neighbour search
def test_knn_l2_and_fk_metric_vs_reference_heap_loop(core, golden, panda_model, cuda):
    g = golden("panda_heap_knn")
    qs = g["samples"]
    q64 = core.to_device_f64(qs)
    q32 = core.pack_f32(q64)
    w = core.SpatialWorld([core.RobotSpec(panda_model, np.eye(4))])
    # L2 in joint space
    F = core.l2_features(q64, q32)
    idx, dist, amb = core.knn_exact(F, F, 10)
    ref_idx, ref_d = g["l2_knn_idx"], g["l2_knn_dist"]
    bad = _np(idx) != ref_idx
    assert np.abs(_np(dist) - ref_d).max() < 1e-12
    # index differences only where the reference's own float64 distances tie within 1e-6
    rows = np.nonzero(bad.any(1))[0]
    for r in rows:
        c = np.nonzero(bad[r])[0]
        assert np.abs(ref_d[r, c] - _np(dist)[r, c]).max() < 1e-6
        assert sorted(_np(idx)[r]) == sorted(ref_idx[r]) or _np(amb)[r] == 1 or \
            abs(ref_d[r, -1] - _np(dist)[r, -1]) < 1e-6
    assert bad.mean() < 0.01
    # FK metric (task_space_pose_distance)
    F = w.fk_features(q64, q32, 0)
    idx, dist, amb = core.knn_exact(F, F, 10)
    ref_idx, ref_d = g["fk_knn_idx"], g["fk_knn_dist"]
    assert np.abs(_np(dist) - ref_d).max() < 1e-10
    assert (_np(idx) != ref_idx).mean() < 0.01
    assert w.metric_groups(0) == ([3, 4, 6, 7, 8], [1.0, 1.0, 2.0, 1.0, 1.0])


def test_knn_vs_scipy_fixture(core, golden, cuda):
    # reference fixtures res/images/{random,halton}_{samples,degree_roadmap}.csv, KDTree k=5
    g = golden("scipy_knn")
    for name in ("random", "halton"):
        s = g[name + "_samples"]
        q64 = core.to_device_f64(s)
        q32 = core.pack_f32(q64)
        F = core.l2_features(q64, q32)
        idx, _, _ = core.knn_exact(F, F, 5)
        pairs = np.stack([np.repeat(np.arange(len(s)), 5), _np(idx).ravel()], 1)
        assert (pairs == g[name + "_degree"]).all()
        off, ridx, rd = core.radius_search(F, F, 0.2)
        off, ridx = _np(off), _np(ridx)
        got = {(i, int(j)) for i in range(len(s)) for j in ridx[off[i]:off[i + 1]]}
        assert got == {tuple(p) for p in g[name + "_distance"]}


@pytest.mark.parametrize("D,n,k", [(7, 20000, 10), (14, 6000, 10), (28, 3000, 16), (2, 5000, 5), (7, 3000, 55)])
def test_knn_large_vs_float64_bruteforce(core, cuda, D, n, k):
    rng = np.random.default_rng(D * 1000 + n)
    q = np.round(rng.uniform(-2.9, 2.9, (n, D)), 2)
    q64 = core.to_device_f64(q)
    q32 = core.pack_f32(q64)
    F = core.l2_features(q64, q32)
    idx, dist, amb = core.knn_exact(F, F, k)
    idx, dist, amb = _np(idx), _np(dist), _np(amb)
    # float64 reference on the GPU box's CPU, blocked
    nq = min(n, 1500)
    sel = rng.choice(n, nq, replace=False)
    d2 = ((q[sel][:, None, :] - q[None, :, :]) ** 2).sum(-1)
    d2[np.arange(nq), sel] = np.inf
    ref_d = np.sqrt(np.sort(d2, 1)[:, :k])
    assert np.abs(dist[sel] - ref_d).max() < 1e-12          # the k smallest float64 distances, exactly
    order = np.lexsort((np.broadcast_to(np.arange(n), d2.shape), d2), axis=1)[:, :k]
    diff = idx[sel] != order
    # an index may differ from (dist, id) order only inside an exact-distance tie group
    for r, c in zip(*np.nonzero(diff)):
        assert abs(d2[r, idx[sel][r, c]] - d2[r, order[r, c]]) < 1e-9      # numpy's vectorised sum differs in the last bits
    assert amb.mean() < 0.05


def test_knn_shards_tile_the_unsharded_search(core, panda_model, cuda):
    """mmmp_knn_shard: the shards of a self-search answer every row exactly once and every answer
    is bit-identical to the unsharded search (culled path at 40 000 rows, exhaustive path at
    3 000, FK metric and L2, shard counts that do not divide the block count)."""
    rng = np.random.default_rng(91)
    w = core.SpatialWorld([core.RobotSpec(panda_model, core.base_matrix((0, 0, 0.01)))], [])
    for n in (3000, 40000):
        q = np.round(rng.uniform(LIMITS[:, 0], LIMITS[:, 1], (n, 7)), 2)
        q64 = core.to_device_f64(q)
        q32 = core.pack_f32(q64)
        for F in (w.fk_features(q64, q32, 0), core.l2_features(q64, q32)):
            ref_i, ref_d = core.knn(F, F, 12, exclude_self=True)
            for shards in (1, 3, 8):
                seen = torch.zeros(n, dtype=torch.int32, device=cuda)
                for s in range(shards):
                    rows, idx, dist = core.knn_shard(F, 12, s, shards)
                    seen[rows.long()] += 1
                    assert (idx == ref_i[rows.long()]).all() and (dist == ref_d[rows.long()]).all()
                assert (seen == 1).all()
            # mmmp_knn_range: uneven cuts (the ones RoadmapBuilder adapts to measured scan times),
            # including an empty slice, tile the search the same way
            cuts = [0.0, 0.07, 0.07, 0.41, 0.93, 1.0]
            seen = torch.zeros(n, dtype=torch.int32, device=cuda)
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                rows, idx, dist = core.knn_range(F, 12, lo, hi)
                if lo == hi:
                    assert rows.numel() == 0
                seen[rows.long()] += 1
                assert (idx == ref_i[rows.long()]).all() and (dist == ref_d[rows.long()]).all()
            assert (seen == 1).all()


def test_empty_and_tiny_inputs(core, panda_model, cuda):
    """Edge cases of every entry point: zero rows, one row, fewer references than k (missing
    neighbours are -1 / inf, as the reference's heap simply holds fewer entries), ragged sizes."""
    from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_SELF
    w = core.SpatialWorld([core.RobotSpec(panda_model, core.base_matrix((0, 0, 0.01)))], [core.plane((0, 0, 1), 0.0)])
    F = CHECK_SELF | CHECK_OBSTACLES
    empty64 = torch.empty((0, 7), dtype=torch.float64, device=cuda)
    empty32 = core.pack_f32(empty64)
    assert empty32.shape == (0, 8)
    assert w.collide_configs(empty32, 0, F).shape == (0,)
    assert core.compact_free(torch.empty((0,), dtype=torch.uint8, device=cuda)).shape == (0,)
    e = torch.empty((0, 2), dtype=torch.int32, device=cuda)
    q = np.round(np.random.default_rng(3).uniform(LIMITS[:, 0], LIMITS[:, 1], (5, 7)), 2)
    q64 = core.to_device_f64(q)
    q32 = core.pack_f32(q64)
    assert w.collide_edges(q32, e, 50, 0.02, 0, F).shape == (0,)
    labels, count = core.connected_components(torch.empty((0, 3), dtype=torch.int32, device=cuda))
    assert labels.shape == (0,) and count == 0
    # 5 references, k = 10: every row has its 4 neighbours, then -1 / inf
    feats = w.fk_features(q64, q32, 0)
    idx, dist, amb = core.knn_exact(feats, feats, 10)
    idx, dist = _np(idx), _np(dist)
    assert (np.sort(idx[:, :4], 1) == np.array([[j for j in range(5) if j != i] for i in range(5)])).all()
    assert (idx[:, 4:] == -1).all() and np.isinf(dist[:, 4:]).all() and np.isfinite(dist[:, :4]).all()
    # a single row: no neighbour at all; its edge slots are reported blocked
    one = w.fk_features(q64[:1].contiguous(), q32[:1].contiguous(), 0)
    idx1, _, _ = core.knn_exact(one, one, 3)
    assert (_np(idx1) == -1).all()
    edges = core.edges_from_knn(idx1)
    assert _np(w.collide_edges(q32[:1].contiguous(), edges, 50, 0.02, 0, F)).tolist() == [1, 1, 1]
    # shards of a tiny set: still every row once
    seen = torch.zeros(5, dtype=torch.int64, device=cuda)
    for s in range(4):
        rows, _, _ = core.knn_shard(feats, 3, s, 4)
        seen += torch.bincount(rows.long(), minlength=5)
    assert (seen == 1).all()


def test_ambiguous_rows_are_searched_again(core, panda_model, cuda):
    """RoadmapBuilder keeps k + 2 fp32 candidates; rows whose rank k cannot be told from the last
    candidate (exact ties here: 3000 rows on a 21 x 21 lattice, hundreds of duplicates) are
    searched again alone with k + 16.  The distances must be the k smallest float64 ones."""
    from mmmp_b200.roadmap import RoadmapBuilder
    rng = np.random.default_rng(17)
    q = np.tile(np.array([[0.0, -0.5, 0.0, -2.0, 0.0, 1.6, 0.8]]), (3000, 1))
    q[:, [1, 3]] += np.round(rng.uniform(-0.1, 0.1, (3000, 2)), 2)
    q64 = core.to_device_f64(q)
    q32 = core.pack_f32(q64)
    w = core.SpatialWorld([core.RobotSpec(panda_model, core.base_matrix((0, 0, 0.01)))], [])
    for metric in ("l2", "fk"):
        out = RoadmapBuilder(w, 0, metric, 10, 20, 0.05, knn_pad=2, single_rank=True).candidates(q64, q32)
        dist = _np(out["dist"])
        if metric == "l2":
            D = np.sqrt(((q[:, None, :] - q[None, :, :]) ** 2).sum(-1))
        else:
            from oracle import chain
            P = chain.panda_positions(q, core.base_matrix((0, 0, 0.01)))
            D = np.stack([np.sqrt(((P[i:i + 1] - P) ** 2).sum(-1)).sum(-1) for i in range(len(q))])
        np.fill_diagonal(D, np.inf)
        ref = np.sort(D, 1)[:, :10]
        assert np.abs(dist - ref).max() < 1e-9
        idx = _np(out["idx"])
        assert (idx != np.arange(len(q))[:, None]).all()            # self never comes back as a neighbour
        assert np.abs(np.take_along_axis(D, idx, 1) - ref).max() < 1e-9   # the ids carry those distances


def test_knn_causal_and_queries_subset(core, cuda):
    rng = np.random.default_rng(5)
    q = np.round(rng.uniform(-3, 3, (4000, 4)), 2)
    q64 = core.to_device_f64(q)
    q32 = core.pack_f32(q64)
    F = core.l2_features(q64, q32)
    # causal: only earlier nodes (incremental planners, distance_prm.py:48)
    idx, dist, _ = core.knn_exact(F, F, 6, causal=True)
    idx, dist = _np(idx), _np(dist)
    assert (idx[0] == -1).all() and (idx[3] >= 0).sum() == 3
    for i in (5, 100, 3999):
        d = np.sqrt(((q[i] - q[:i]) ** 2).sum(-1))
        o = np.lexsort((np.arange(i), d))[:6]
        assert np.allclose(dist[i][:len(o)], d[o], atol=1e-12)
        assert (np.sort(idx[i][:len(o)]) == np.sort(o)).all() or np.allclose(d[idx[i][:len(o)]], d[o])
    # a shard of queries with global ids (multi-GPU sharding): rows 1000..1499
    idx2, dist2, _ = core.knn_exact(F, F.rows(1000, 1500), 6, query_id0=1000)
    full, dfull, _ = core.knn_exact(F, F, 6)
    assert (_np(idx2) == _np(full)[1000:1500]).all()


def test_radius_vs_reference_distance_lists(core, golden, panda_model, cuda):
    """connect_nearest_neighbors_distance candidate sets (0 < d < maxdist), FK metric and L2:
    with an always-free edge rule the reference's adjacency is exactly the neighbour sets."""
    g = golden("panda_heap_knn")
    qs = g["samples"]
    q64 = core.to_device_f64(qs)
    q32 = core.pack_f32(q64)
    w = core.SpatialWorld([core.RobotSpec(panda_model, np.eye(4))])
    F = core.l2_features(q64, q32)
    off, idx, d = core.radius_search(F, F, 1.6, strict=True, exclude_zero=True)
    off, idx, d = _np(off), _np(idx), _np(d)
    ref = np.sqrt(((qs[:, None] - qs[None]) ** 2).sum(-1))
    for i in range(len(qs)):
        want = np.nonzero((ref[i] > 0) & (ref[i] < 1.6))[0]
        assert (idx[off[i]:off[i + 1]] == want).all()
    F = w.fk_features(q64, q32, 0)
    off, idx, d = core.radius_search(F, F, 2.0, strict=True, exclude_zero=True)
    off, idx, d = _np(off), _np(idx), _np(d)
    from oracle import chain
    for i in range(0, len(qs), 7):
        dd = chain.panda_pose_distance(np.repeat(qs[i:i + 1], len(qs), 0), qs)
        want = np.nonzero((dd > 0) & (dd < 2.0))[0]
        assert (idx[off[i]:off[i + 1]] == want).all()
        assert np.abs(d[off[i]:off[i + 1]] - dd[want]).max() < 1e-12


# --------------------------------------------------------------------- host-buffer call
def test_roadmap_candidates_host_end_to_end(core, panda_model, cuda):
    from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_SELF
    from oracle import chain
    from oracle.spheres import SphereChecker
    rng = np.random.default_rng(21)
    base = core.base_matrix((0, 0, 0.01))
    obs = _scene(core)
    w = core.SpatialWorld([core.RobotSpec(panda_model, base)], obs)
    ck = SphereChecker([panda_model], [base], obs)
    q = np.round(rng.uniform(LIMITS[:, 0], LIMITS[:, 1], (3000, 7)), 2)
    for metric in (core.METRIC_L2, core.METRIC_GROUP_L2_SUM):
        out = w.roadmap_candidates_host(q, 0, metric, k=8, n_steps=50, step=0.02)
        free_ref = ~ck.config_collision(q, 0)
        src = out["node_src"]
        assert abs(len(src) - free_ref.sum()) <= 3
        keep = np.nonzero(free_ref)[0]
        if len(src) != len(keep) or (src != keep).any():
            continue_ok = len(set(src) ^ set(keep)) <= 3
            assert continue_ok
            continue
        nodes = q[src]
        if metric == core.METRIC_L2:
            D = np.sqrt(((nodes[:, None] - nodes[None]) ** 2).sum(-1))
        else:
            P = chain.panda_positions(nodes, base)
            D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1)).sum(-1)
        np.fill_diagonal(D, np.inf)
        ref_d = np.sort(D, 1)[:, :8]
        assert np.abs(out["nbr_dist"] - ref_d).max() < 1e-9
        nb = out["nbr_idx"]
        sel = rng.choice(len(nodes), 150, replace=False)
        qa = np.repeat(nodes[sel], 8, 0)
        qb = nodes[nb[sel].ravel()]
        eref = ~ck.edge_collision(qa, qb, 50, 0.02, 0)
        assert (out["edge_free"][sel].ravel().astype(bool) != eref).sum() <= 3
        assert out["timings_ms"][7] > 0


# --------------------------------------------------------------------- (8f-4) batched IK
def test_ik_kernel_vs_float64_oracle(core, panda_model, cuda):
    """mmmp_ik_dls against the float64 restatement of the same iteration (oracle/ik.py): same
    convergence verdicts, residuals below tolerance, joint values equal to 2e-3 rad where the
    iteration is contractive; position-only and full-pose targets; shared start row."""
    from oracle import chain, ik, spheres
    base = core.base_matrix((0.1, -0.2, 0.01), (0, 0, 0.38268343, 0.92387953))
    w = core.SpatialWorld([core.RobotSpec(panda_model, base)], [])
    rng = np.random.default_rng(31)
    n = 3000
    q_true = rng.uniform(LIMITS[:, 0] * 0.6, LIMITS[:, 1] * 0.6, (n, 7))
    q_true[:, 3] = rng.uniform(-2.4, -0.8, n)
    F = chain.chain_frames(spheres.model_origins(panda_model), q_true, base)[:, -1]
    q0 = np.clip(q_true + rng.normal(0, 0.2, q_true.shape), LIMITS[:, 0], LIMITS[:, 1])
    R = F[:, :3, :3]
    qw = np.sqrt(np.maximum(1 + R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2], 1e-12)) / 2
    quat = np.stack([(R[:, 2, 1] - R[:, 1, 2]) / (4 * qw), (R[:, 0, 2] - R[:, 2, 0]) / (4 * qw),
                     (R[:, 1, 0] - R[:, 0, 1]) / (4 * qw), qw], 1)
    tp = core.to_device_f64(np.ascontiguousarray(F[:, :3, 3]))
    for tq_np in (None, quat):
        tq = None if tq_np is None else core.to_device_f64(tq_np)
        q, conv, res = w.ik(tp, tq, core.to_device_f64(q0), 0, max_iters=150)
        q, conv, res = q.cpu().numpy(), conv.cpu().numpy().astype(bool), res.cpu().numpy()
        q_ref, ok_ref, res_ref = ik.ik_dls(panda_model, base, F[:, :3, 3], tq_np, q0, max_iters=150)
        assert conv.mean() > 0.97 and (conv == ok_ref).mean() > 0.995
        both = conv & ok_ref
        assert res[both, 0].max() < 1e-4 and res[both, 1].max() < 1e-3
        # the tool pose of the answer, recomputed in float64, meets the target
        e, _ = ik.task_error(panda_model, base, q[both], F[both, :3, 3], None if tq_np is None else tq_np[both])
        assert np.linalg.norm(e[:, :3], axis=1).max() < 2e-4
        assert np.abs(q[both] - q_ref[both]).max() < (2e-3 if tq_np is not None else 5e-2)
        assert ((q >= LIMITS[:, 0] - 1e-6) & (q <= LIMITS[:, 1] + 1e-6)).all()
    # one start row for all problems (ld0 = 0), unreachable targets are reported, not faked
    far = np.tile([[3.0, 0.0, 0.5]], (5, 1))
    q, conv, res = w.ik(core.to_device_f64(far), None, core.to_device_f64(q0[:1]), 0, max_iters=50)
    assert not conv.cpu().numpy().any() and res.cpu().numpy()[:, 0].min() > 1.0


def test_rerank_lanes_per_row_equals_the_sequential_rerank(core, cuda):
    """mmmp_knn_refine ranks a row's candidates with one lane per candidate; rows with an exact float64
    tie across rank k go through the sequential rule (the reference's bounded heap keeps the LARGER ids,
    decoupled_prm.py:490-492).  Every output must equal the one-thread-per-row kernel's, on data made of
    repeated points (ties everywhere), with empty candidate slots and with more than 16 candidates."""
    import os
    rng = np.random.default_rng(31)
    base = np.round(rng.uniform(-1, 1, (500, 7)), 1)
    q = base[rng.integers(0, 500, 4000)]                          # every point ~8 times: exact ties
    feats = core.l2_features(core.to_device_f64(q))
    for kc, k in ((12, 10), (12, 12), (26, 10), (16, 3)):
        ci, cd = core.knn(feats, feats, kc, exclude_self=True)
        ci[5, 7:] = -1                                             # a short row
        ci[6, :] = -1                                              # an empty row
        got = core.knn_refine(feats.x64, feats.x64, feats.m64, ci, cd, k)
        os.environ["MMMP_REFINE_SERIAL"] = "1"
        try:
            want = core.knn_refine(feats.x64, feats.x64, feats.m64, ci, cd, k)
        finally:
            del os.environ["MMMP_REFINE_SERIAL"]
        for a, b in zip(got, want):
            assert torch.equal(a, b), (kc, k)
    # the data exercises the tie rule: float64 brute force on some rows, ties across rank 10
    D = np.sqrt(((q[:200, None, :] - q[None, :, :]) ** 2).sum(-1))
    D[np.arange(200), np.arange(200)] = np.inf
    D.sort(1)
    assert (D[:, 10] == D[:, 9]).sum() > 50
