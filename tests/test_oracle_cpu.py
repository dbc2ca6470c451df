"""CPU suite: the oracle restatements against the reference's own outputs
(tests/golden/*.npz, written by tools/gen_golden.py from the reference itself) and
against the reference's fixtures / known answers."""
import numpy as np
import pytest

from oracle import chain, knn as oknn, planar, sampling


def _planar_world(g, n_arms=2):
    rects = [{"type": "rect", "x0": r[0], "y0": r[1], "x1": r[2], "y1": r[3]} for r in g["w2_rects"]]
    circ = [{"type": "circle", "cx": c[0], "cy": c[1], "r": c[2]} for c in g["w2_circles"]]
    arms = [{"links": list(g["w2_links"][i]), "base": tuple(g["w2_bases"][i])} for i in range(n_arms)]
    return {"arms": arms, "bounds": (-10.0, 10.0, -10.0, 10.0), "obstacles": rects + circ, "joint_link_radius": 0.5}


def test_planar_fk_known_answer():
    # reference tests/robots/test_planar_arm.py:19-26
    pos = planar.fk([1.0, 1.0], (0.0, 0.0), [np.pi / 2, np.pi / 2])
    np.testing.assert_almost_equal(np.array(pos), np.array([[0, 1], [-1, 1]]), decimal=5)


def test_planar_predicates_match_reference(golden):
    g = golden("planar_predicates")
    w = _planar_world(g)
    q = g["coupled_q"]
    got = np.array([not planar.config_in_collision(w, s) for s in q[:1500]])
    assert (got == g["coupled_free"][:1500]).all()
    rr = np.array([planar.config_in_collision(w, s, bounds=False, self_=False, obstacles=False) for s in q[:1500]])
    assert (rr == g["rr_hit"]).all()
    q2 = g["dec_q"]
    for arm, key in ((0, "dec_free_r0"), (1, "dec_free_r1")):
        got = np.array([not planar.config_in_collision(w, s, arm=arm, self_=False, robots=False) for s in q2[:800]])
        assert (got == g[key][:800]).all()


def test_planar_edges_match_reference(golden):
    g = golden("planar_predicates")
    w = _planar_world(g)
    q = g["coupled_q"]
    got = np.array([not planar.edge_in_collision(w, a, b, 0.05) for a, b in zip(q[:150], q[400:550])])
    assert (got == g["coupled_edge_free"][:150]).all()
    got = np.array([not planar.edge_in_collision(w, a, b, 0.05) for a, b in zip(q[:150], g["coupled_edge2_b"][:150])])
    assert (got == g["coupled_edge2_free"][:150]).all()


def test_panda_fk_matches_reference(golden):
    g = golden("panda_fk")
    for tag in ("id", "b1", "b2"):
        b = g[f"fk_{tag}_base"]
        base = chain.base_matrix(b[:3], b[3:])
        pos = chain.panda_positions(g[f"fk_{tag}_q"], base)
        assert np.abs(pos - g[f"fk_{tag}_pos"]).max() < 1e-12
        T = chain.panda_frames(g[f"fk_{tag}_q"], base)[:, 7]
        assert np.abs(T - g[f"fk_{tag}_T"]).max() < 1e-12
    q = g["fk_id_q"]
    d = chain.panda_pose_distance(q[:200], q[200:])
    assert np.abs(d - g["pose_dist"]).max() < 1e-12
    # SURVEY 8c known answers: flange at q_ref, and the pair distance
    q_ref = [0, -np.pi / 4, 0, -3 * np.pi / 4, 0, np.pi / 2, np.pi / 4]
    assert np.allclose(chain.panda_positions([q_ref])[0, 7], [0.306891, 0.0, 0.590282], atol=1e-6)
    assert abs(chain.panda_pose_distance([q_ref], [[.1, -.5, .2, -2, .3, 1.5, .7]])[0] - g["known_dist"][0]) < 1e-12
    assert np.allclose(chain.panda_positions([[0] * 7])[0, 7], [0.088, 0, 0.926], atol=1e-12)


def test_generic_chain_equals_dh_chain(golden):
    from oracle.spheres import model_origins
    import json, os
    model = json.load(open(os.path.join(os.path.dirname(__file__), "..", "mmmp_b200", "geometry", "panda_spheres.json")))
    q = golden("panda_fk")["fk_id_q"][:50]
    fr = chain.chain_frames(model_origins(model), q, np.eye(4))
    assert np.abs(fr[:, :, :3, 3] - chain.panda_positions(q)).max() < 1e-12


def test_heap_knn_and_greedy_match_reference(golden):
    g = golden("panda_heap_knn")
    n12 = np.array([[1, 1], [1, 2], [2, 1], [10, 10], [11, 10], [10, 11], [20, 20], [21, 20], [20, 21], [30, 30],
                    [30, 31], [31, 30]], float)
    D = np.linalg.norm(n12[:, None] - n12[None], axis=-1)
    c = oknn.heap_knn(lambda i, j: D[i, j], 12, 5)
    adj = oknn.greedy_degree(c, lambda i, j: True, 3)
    assert [len(a) for a in adj] == list(g["n12_degree_deg"])
    assert sum(adj, []) == list(g["n12_degree_adj"])
    # production method output on the 12-node case
    assert adj[4] == [5, 6] and adj[8] == [7, 9, 5] and adj[11] == [10]
    qs = g["samples"]
    for metric in ("l2", "fk"):
        if metric == "l2":
            Dm = np.sqrt(((qs[:, None] - qs[None]) ** 2).sum(-1))
        else:
            Dm = np.stack([chain.panda_pose_distance(np.repeat(qs[i:i + 1], len(qs), 0), qs) for i in range(len(qs))])
        idx, dist = oknn.brute_knn(Dm, 10)
        ref_idx, ref_d = g[f"{metric}_knn_idx"], g[f"{metric}_knn_dist"]
        bad = idx != ref_idx
        # only float64 last-bit ties may differ
        assert bad.mean() < 0.01
        assert np.abs(dist - ref_d).max() < 1e-9


def test_shared_verdicts_build_the_reference_roadmap(golden):
    """mmmp_collide_knn_edges checks a pair of mutual candidates once (low id -> high id).  Restated in
    oracle.knn.shared_verdict_table and fed to the accept pass, it must build exactly the adjacency the
    reference's production connect_nearest_neighbors_degree built on the golden samples (its edge rule is
    the golden's; the reference loop asked for more checks than there are distinct pairs)."""
    g = golden("panda_heap_knn")
    qs = g["samples"]

    def ef(a, b):                                              # tools/gen_golden.py ef(): the golden's edge rule
        return (int(round((qs[a].sum() + qs[b].sum()) * 100)) % 7) != 0

    for metric in ("fk", "l2"):
        idx, dist = g[f"{metric}_knn_idx"], g[f"{metric}_knn_dist"]
        free, calls = oknn.shared_verdict_table(idx, ef)
        assert len(set(calls)) == len(calls)
        pairs = {(min(i, int(j)), max(i, int(j))) for i in range(len(qs)) for j in idx[i]}
        assert len({(min(a, b), max(a, b)) for a, b in calls}) == len(pairs) == len(calls)
        mutual = sum(1 for i in range(len(qs)) for j in idx[i] if i in idx[int(j)])
        assert 0.3 < mutual / idx.size < 0.95 and len(calls) == idx.size - mutual // 2
        # every mutual pair is walked from its lower id: the first check of the reference's loop
        assert all(a < b for a, b in calls if b in idx[a] and a in idx[b])
        table = {(i, int(j)): bool(free[i, s]) for i in range(len(qs)) for s, j in enumerate(idx[i])}
        cands = [[(dist[i, s], int(idx[i, s])) for s in range(idx.shape[1])] for i in range(len(qs))]
        asked = []

        def from_table(i, j):
            asked.append((i, j))
            return table[(i, j)]

        adj = oknn.greedy_degree(cands, from_table, 6)
        assert [len(a) for a in adj] == list(g[f"{metric}_degree_deg"])
        assert sum(adj, []) == list(g[f"{metric}_degree_adj"])
        # the reference's loop is lazy (it stops at k2 = 6 accepted edges) AND repeats itself: some pairs are
        # asked for from both ends; the table holds one verdict for both
        twice = len(asked) - len({(min(a, b), max(a, b)) for a, b in asked})
        assert twice > 20 and len(asked) <= idx.size


def test_scipy_fixture_knn(golden):
    # reference fixtures res/images/{random,halton}_{samples,degree_roadmap}.csv (k=5)
    g = golden("scipy_knn")
    for name in ("random", "halton"):
        s = g[name + "_samples"]
        D = np.sqrt(((s[:, None] - s[None]) ** 2).sum(-1))
        idx, _ = oknn.brute_knn(D, 5)
        pairs = np.stack([np.repeat(np.arange(len(s)), 5), idx.ravel()], 1)
        assert (pairs == g[name + "_degree"]).all()
        # radius 0.2 fixture: same SET of directed pairs (cKDTree order inside a ball is unspecified)
        got = {(i, j) for i in range(len(s)) for j in range(len(s)) if i != j and D[i, j] <= 0.2}
        assert got == {tuple(p) for p in g[name + "_distance"]}


def test_philox_known_answer_and_rounding():
    # Random123 known-answer test vector for philox4x32-10 (counter = key = 0)
    r = sampling.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(x[0]) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = sampling.philox4x32_10([0xffffffff], [0xffffffff], [0xffffffff], [0xffffffff], 0xffffffff, 0xffffffff)
    assert [int(x[0]) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    s = sampling.sample_uniform([-1, 0], [1, 3], 1000, seed=7)
    assert ((s >= [-1, 0]) & (s <= [1, 3])).all()
    assert np.abs(s * 100 - np.rint(s * 100)).max() < 1e-9
    assert (s == np.round(s, 2)).all()


def test_hull_oracle_known_answers():
    from oracle import hull
    s = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], float)
    assert hull.distance(s, s + [3, 0, 0]) == pytest.approx(2.0)
    assert hull.distance(s, s + [1.5, 1.5, 0]) == pytest.approx(np.sqrt(0.5))
    assert hull.distance(s, s + [2, 2, 2]) == pytest.approx(np.sqrt(3.0))
    assert hull.distance(s, s + [0.5, 0.5, 0.5]) == 0.0
    assert hull.distance(s, s + [1, 0, 0]) == 0.0          # touching = collision (distance <= 0)
    assert hull.distance(s[:1], s[:1] + [0, 0, 5], 1.0, 1.5) == pytest.approx(2.5)   # two spheres
    rng = np.random.default_rng(0)
    for _ in range(50):   # random tetrahedra vs brute-force sampled distance upper bound
        a, b = rng.normal(size=(4, 3)), rng.normal(size=(4, 3)) + [4, 0, 0]
        d = hull.distance(a, b)
        w = rng.dirichlet(np.ones(4), size=(2, 4000))
        ub = np.linalg.norm((w[0] @ a)[:, None] - (w[1] @ b)[None], axis=-1).min()
        assert d <= ub + 1e-9 and d > ub - 0.25
    # a pose where GJK builds a flat tetrahedron (regression: sign noise read as containment):
    # the distance must not depend on the frame the pair is posed in
    from mmmp_b200 import core
    q = np.array([[0.31, 0.35, -0.84, -1.26, 1.33, 0.33, 0.21]])
    P = hull.link_poses(q, core.base_matrix((0, 0, 0.01)))
    H = hull.hulls()
    TA, TB = P[:, hull.LINK_FRAME[5]], P[:, hull.LINK_FRAME[7]]
    d_world = hull.posed_distance(H[5], TA, H[7], TB)[0]
    d_local = hull.posed_distance(H[5], np.eye(4)[None], H[7], (np.linalg.inv(TA[0]) @ TB[0])[None])[0]
    assert d_world == pytest.approx(d_local, abs=1e-9) and d_world == pytest.approx(0.01106909, abs=1e-7)
    assert not hull.PandaHullChecker([core.base_matrix((0, 0, 0.01))]).self_collision(q)[0]


def test_hull_oracle_vs_linear_programming():
    """An independent algorithm for the same question (pybullet itself is absent on the build and the
    GPU box: profiles/r02_pybullet_probe.txt): two convex hulls intersect iff the LP
    { lambda, mu >= 0, sum lambda = sum mu = 1, A^T lambda = B^T mu } is feasible (scipy HiGHS), and
    their distance is the optimum of the QP min |A^T lambda - B^T mu| (SLSQP).  The GJK oracle must
    give the same verdicts and distances on the real Panda link hulls in random relative poses."""
    from scipy.optimize import linprog, minimize
    from mmmp_b200 import core
    from oracle import hull
    H = hull.hulls()
    rng = np.random.default_rng(12)
    pairs = [(1, 5), (2, 5), (5, 7), (2, 6), (0, 7), (1, 6), (3, 8), (4, 9)]
    n_hit = n_checked = 0
    for n in range(48):
        la, lb = pairs[n % len(pairs)]
        # link lb in a random pose around link la: random rotation, centroids 0 .. 0.3 m apart
        Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        Q *= np.sign(np.linalg.det(Q))
        u = rng.normal(size=3)
        TA, TB = np.eye(4), np.eye(4)
        TB[:3, :3] = Q
        TB[:3, 3] = H[la].mean(0) + u / np.linalg.norm(u) * rng.uniform(0.0, 0.3) - Q @ H[lb].mean(0)
        A = H[la] @ TA[:3, :3].T + TA[:3, 3]
        B = H[lb] @ TB[:3, :3].T + TB[:3, 3]
        d = hull.posed_distance(H[la], TA[None], H[lb], TB[None])[0]
        na, nb = len(A), len(B)
        Aeq = np.zeros((5, na + nb))
        Aeq[:3, :na], Aeq[:3, na:] = A.T, -B.T
        Aeq[3, :na] = 1
        Aeq[4, na:] = 1
        lp = linprog(np.zeros(na + nb), A_eq=Aeq, b_eq=[0, 0, 0, 1, 1], bounds=(0, None), method="highs")
        feasible = lp.status == 0
        if abs(d) > 1e-6 or feasible:                      # (a grazing contact may be either for the LP)
            assert feasible == (d <= 0.0), (n, la, lb, d, lp.status)
        n_hit += int(d <= 0.0)
        if d > 1e-3 and n_checked < 5:                    # distance by an independent QP on a few separated pairs
            x0 = np.concatenate([np.full(na, 1 / na), np.full(nb, 1 / nb)])
            res = minimize(lambda x: ((x[:na] @ A - x[na:] @ B) ** 2).sum(), x0, method="SLSQP",
                           jac=lambda x: np.concatenate([2 * A @ (x[:na] @ A - x[na:] @ B), -2 * B @ (x[:na] @ A - x[na:] @ B)]),
                           bounds=[(0, 1)] * (na + nb), options={"maxiter": 400, "ftol": 1e-14},
                           constraints=[{"type": "eq", "fun": lambda x: x[:na].sum() - 1},
                                        {"type": "eq", "fun": lambda x: x[na:].sum() - 1}])
            assert np.sqrt(res.fun) == pytest.approx(d, abs=2e-4), (n, la, lb)
            n_checked += 1
    assert 8 <= n_hit <= 40 and n_checked >= 4


def test_task_space_position_sampler_matches_reference(golden):
    """oracle.ik.task_space_positions restates sample_task_space_position (decoupled_prm.py:346-369);
    golden = the reference's own function under a seeded numpy stream + the draws it consumed."""
    from oracle import ik
    g = golden("task_space_sampler")
    for c in range(3):
        d = g[f"draws{c}"]
        got = ik.task_space_positions(g[f"start{c}"], g[f"goal{c}"], d[:, 0], d[:, 1], d[:, 2])
        assert np.abs(got - g[f"pos{c}"]).max() < 1e-12


def test_ik_oracle_reaches_reachable_poses():
    """The float64 DLS restatement converges on poses generated by FK (known to be reachable)."""
    from mmmp_b200 import core
    from oracle import ik, spheres
    model = core.load_model("panda_spheres")
    base = core.base_matrix((0.1, -0.2, 0.01), (0, 0, 0.38268343, 0.92387953))
    lim = np.array(model["joint_limits"])
    rng = np.random.default_rng(3)
    q_true = rng.uniform(lim[:, 0] * 0.6, lim[:, 1] * 0.6, (40, 7))
    q_true[:, 3] = rng.uniform(-2.4, -0.8, 40)
    F = chain.chain_frames(spheres.model_origins(model), q_true, base)[:, -1]
    q0 = np.clip(q_true + rng.normal(0, 0.25, q_true.shape), lim[:, 0], lim[:, 1])
    q, ok, res = ik.ik_dls(model, base, F[:, :3, 3], None, q0)
    assert ok.all() and res[:, 0].max() < 1e-4
    # with orientation: target quaternion from the rotation matrices
    R = F[:, :3, :3]
    w = np.sqrt(np.maximum(1 + R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2], 1e-12)) / 2
    quat = np.stack([(R[:, 2, 1] - R[:, 1, 2]) / (4 * w), (R[:, 0, 2] - R[:, 2, 0]) / (4 * w),
                     (R[:, 1, 0] - R[:, 0, 1]) / (4 * w), w], 1)
    assert np.abs(ik.quat_to_rot(quat) - R).max() < 1e-9
    q, ok, res = ik.ik_dls(model, base, F[:, :3, 3], quat, q0, max_iters=200)
    assert ok.mean() > 0.9 and res[ok].max() < 1e-3
    assert ((q >= lim[:, 0] - 1e-12) & (q <= lim[:, 1] + 1e-12)).all()
