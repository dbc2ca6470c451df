"""GPU parity tests of the planner layer (the reference's plugin API for the hot path):
roadmaps built by the mmmp_b200 planner classes from FIXED samples must equal, list by list,
the roadmaps the reference's own classes built from the same samples
(tests/golden/planar_roadmaps.npz, panda_heap_knn.npz -- written by tools/gen_golden.py from
the reference itself), and planned CBS-PRM paths must be collision free under the hull
oracle (the restatement of the reference checker)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LIMITS = np.array([(-2.8973, 2.8973), (-1.7628, 1.7628), (-2.8973, 2.8973), (-3.0718, -0.0698),
                   (-2.8973, 2.8973), (-0.0175, 3.7525), (-2.8973, 2.8973)])


def _adj(roadmap, n):
    deg = np.array([len(roadmap[i]) for i in range(n)])
    flat = np.concatenate([np.array(roadmap[i], dtype=np.int64) for i in range(n)] + [np.zeros(0, np.int64)])
    return deg, flat


@pytest.fixture(scope="module")
def planar_envs(cuda, golden):
    from mmmp_b200.planners.prm.planar.env import Environment
    from mmmp_b200.robots.planar_arm import PlanarArm
    from mmmp_b200.utils.shapes import Circle, Rectangle
    g = golden("planar_predicates")
    arm3 = PlanarArm(np.array([2.5, 2.0, 2.2]), np.array([0.5, -1.0]))
    env3 = Environment((20, 20), [{"name": "a", "start": np.array([0.1, 0.2, 0.3]), "goal": np.array([0.3, 0.2, 0.1]),
                                   "model": arm3}],
                       [Rectangle((3.0, 1.0), 2.0, 2.0), Circle((-2.5, 2.5), 1.0)])
    bases = g["w2_bases"]
    arms = [PlanarArm(np.array([3.0, 3.0]), np.array(b)) for b in bases]
    starts = [np.array([1.2, 0.4]), np.array([2.0, -0.4])]
    agents = [{"name": f"a{i}", "start": starts[i], "goal": starts[i] + 0.3, "model": arms[i]} for i in range(2)]
    env2 = Environment((20, 20), agents, [Rectangle((-1.5, -2.0), 3.0, 1.2), Rectangle((2.1, 3.3), 1.7, 2.9),
                                          Circle((-3.0, 1.5), 1.1)])
    return env3, env2


def test_planar_environment_matches_reference_geometry(planar_envs, golden):
    from mmmp_b200.planners.prm.planar.env import obstacle_desc
    g = golden("planar_predicates")
    _, env2 = planar_envs
    rects = [obstacle_desc(o) for o in env2.obstacles if hasattr(o, "get_bbox")]
    got = np.array([[r["x0"], r["y0"], r["x1"], r["y1"]] for r in rects])
    assert (got == g["w2_rects"]).all()          # the bbox arithmetic of env.py:170 reproduced bit for bit


def test_planar_single_arm_roadmaps_equal_reference(planar_envs, golden):
    from mmmp_b200.planners.prm.planar.single_arm.degree_prm import KDTreeKNNPRM, KNNPRM
    from mmmp_b200.planners.prm.planar.single_arm.distance_prm import DistancePRM, KDTreeDistancePRM
    env3, _ = planar_envs
    g = golden("planar_roadmaps")
    samples = g["single_samples"]
    n = len(samples)

    def build(cls, **kw):
        p = cls.__new__(cls)
        from mmmp_b200.planners.prm.planar.single_arm.prm import PRM
        PRM.__init__(p, env3, 0.05)
        for k_, v_ in kw.items():
            setattr(p, k_, v_)
        return p

    p = build(DistancePRM, max_edge_len=6.0)
    p.add_n_samples(n, samples=samples)
    deg, flat = _adj(p.roadmap, n)
    assert (deg == g["distance_deg"]).all() and (flat == g["distance_adj"]).all()

    p = build(KNNPRM, max_edge_len=6.0, n_knn=4)
    p.add_n_samples(n, samples=samples)
    deg, flat = _adj(p.roadmap, n)
    assert (deg == g["knn_deg"]).all() and (flat == g["knn_adj"]).all()

    p = build(KDTreeKNNPRM, max_edge_len=1.2, n_knn=4)
    p.add_n_samples(n, samples=samples)
    p.update_roadmap()
    deg, flat = _adj(p.roadmap, n)
    assert (deg == g["kd_knn_deg"]).all() and (flat == g["kd_knn_adj"]).all()

    p = build(KDTreeDistancePRM, max_edge_len=1.2)
    p.add_n_samples(n, samples=samples)
    p.update_roadmap()
    deg, flat = _adj(p.roadmap, n)
    assert (deg == g["kd_distance_deg"]).all()
    # cKDTree's order inside a ball is unspecified: compare the neighbour SETS per node
    off = np.concatenate([[0], np.cumsum(deg)])
    ref = g["kd_distance_adj"]
    assert all(sorted(flat[off[i]:off[i + 1]]) == sorted(ref[off[i]:off[i + 1]]) for i in range(n))


def test_planar_multi_arm_roadmaps_equal_reference(planar_envs, golden):
    from mmmp_b200.planners.prm.planar.multi_arm.coupled.coupled_prm import CoupledPRM
    from mmmp_b200.planners.prm.planar.multi_arm.coupled.degree_coupled_prm import KDTreeKNNCoupledPRM
    from mmmp_b200.planners.prm.planar.multi_arm.coupled.distance_coupled_prm import DistanceCoupledPRM
    from mmmp_b200.planners.prm.planar.multi_arm.decoupled.cbs_prm import CBSPRM
    from mmmp_b200.planners.prm.planar.multi_arm.decoupled.decoupled_prm import DecoupledPRM
    _, env2 = planar_envs
    g = golden("planar_roadmaps")
    samples = g["coupled_samples"]
    n = len(samples)
    p = DistanceCoupledPRM.__new__(DistanceCoupledPRM)
    CoupledPRM.__init__(p, env2)
    p.max_edge_len = 1.6
    p.add_n_samples(n, samples=samples)
    deg, flat = _adj(p.roadmap, n)
    assert (deg == g["coupled_distance_deg"]).all() and (flat == g["coupled_distance_adj"]).all()

    p = KDTreeKNNCoupledPRM.__new__(KDTreeKNNCoupledPRM)
    CoupledPRM.__init__(p, env2)
    p.max_edge_len, p.n_knn = 1.6, 5
    p.add_n_samples(n, samples=samples)
    p.update_roadmap()
    deg, flat = _adj(p.roadmap, n)
    assert (deg == g["coupled_kdknn_deg"]).all() and (flat == g["coupled_kdknn_adj"]).all()

    s1 = g["cbs_samples_r1"]
    p = CBSPRM.__new__(CBSPRM)
    DecoupledPRM.__init__(p, env2)
    p.max_edge_len = 0.9
    p.add_n_samples(len(s1), 1, samples=s1)
    deg, flat = _adj(p.roadmaps[1], len(s1))
    assert (deg == g["cbs_deg_r1"]).all() and (flat == g["cbs_adj_r1"]).all()


def test_planar_planner_builds_from_its_own_sampler(planar_envs):
    from mmmp_b200.planners.prm.planar.multi_arm.decoupled.cbs_prm import CBSPRM
    _, env2 = planar_envs
    p = CBSPRM(env2, max_edge_len=1.0, build_type='n_nodes', n_nodes=300, seed=3)
    assert p.compute_combined_nr_nodes() == 600
    assert p.compute_combined_avg_degree() > 2.0
    for r in (0, 1):
        for node in p.node_lists[r][:50]:
            assert p.is_collision_free_sample(node.q, r)
        # symmetric adjacency
        rm = p.roadmaps[r]
        assert all(i in rm[j] for i in rm for j in rm[i])


# ----------------------------------------------------------------------------- Panda
def _panda_env(n_arms=1, obstacles=None):
    from mmmp_b200 import core
    from mmmp_b200.planners.prm.pybullet.env import Environment
    from mmmp_b200.robots.panda import Panda
    start = (0., 0., 0., round(-3 * np.pi / 4, 3), 0., round(3 * np.pi / 4, 3), round(np.pi / 4, 3))
    goal = (0., 0., 0., round(-np.pi / 2 + np.pi / 16, 3), 0., round(np.pi / 2 - np.pi / 16, 3), round(np.pi / 4, 3))
    if n_arms == 1:
        pandas = [Panda(base_position=(0, 0, 0.01))]
    else:   # sims/cbs_prm.py:27-28
        pandas = [Panda(base_position=(0.7, 0, 0.01), base_orientation=(0, 0, 1, 0)), Panda(base_position=(-0.7, 0, 0.01))]
    agents = [{"name": f"agent{i + 1}", "start": start, "goal": goal, "model": p} for i, p in enumerate(pandas)]
    return Environment(agents, [core.plane((0, 0, 1), 0.0)] if obstacles is None else obstacles)


def test_panda_robot_interface(cuda):
    from mmmp_b200.robots.panda import Panda
    p = Panda(base_position=(0.1, 0.2, 0.01), base_orientation=(0, 0, 1, 0))
    assert p.arm_dimension == 7 and len(p.gripper_joints) == 2 and len(p.c_space) == 9
    assert p.arm_c_space[3].upper == pytest.approx(-0.0698)
    p.set_arm_pose([0.1] * 7)
    assert p.get_arm_pose() == (0.1,) * 7
    q_ref = [0, -np.pi / 4, 0, -3 * np.pi / 4, 0, np.pi / 2, np.pi / 4]
    p0 = Panda()
    assert np.allclose(p0.position_from_fk(q_ref), [0.306891, 0.0, 0.590282], atol=1e-6)
    assert p0.task_space_pose_distance(q_ref, [.1, -.5, .2, -2, .3, 1.5, .7]) == pytest.approx(0.8277714713081394, abs=1e-12)
    with pytest.raises(ValueError):
        from mmmp_b200.robots.robot import Robot
        Robot("something.sdf")


def test_decoupled_connect_methods_equal_reference(cuda, golden):
    """connect_nearest_neighbors_degree / _distance on the golden 300 samples with the golden edge
    rule: edge_dicts must equal what the reference's production methods produced."""
    from mmmp_b200.planners.prm.pybullet.decoupled.decoupled_prm import DecoupledPRM
    g = golden("panda_heap_knn")
    qs = g["samples"]
    env = _panda_env(1)

    class Probe(DecoupledPRM):
        def _edges_hit(self, r_index, q32, edges):
            e = edges.cpu().numpy()
            q = np.array([n.q for n in self.node_lists[r_index]])
            hit = [(int(round((q[a].sum() + q[b].sum()) * 100)) % 7) == 0 for a, b in e]   # tools/gen_golden.py ef()
            return torch.tensor(hit, dtype=torch.uint8, device=q32.device)

        def _candidate_verdicts(self, r_index, q32, idx):
            from mmmp_b200 import core
            return (self._edges_hit(r_index, q32, core.edges_from_knn(idx)) == 0).to(torch.uint8).reshape(idx.shape)

    p = Probe.__new__(Probe)
    p.env, p.agents, p.robot_models, p.obstacles = env, env.agents, env.robot_models, env.obstacles
    p.r_ids = [m.r_id for m in env.robot_models]
    p.c_spaces = p.create_c_spaces(env.robot_models)
    p.k1, p.k2, p.maxdist, p.local_step, p.time_step = 10, 6, 2.0, 0.02, 0.1
    for mode in ("degree", "distance"):
        p.node_lists, p.edge_dicts, p.node_id_q_dicts, p.node_q_id_dicts = [[]], [{}], [{}], [{}]
        p._append_nodes(0, qs)
        getattr(p, "connect_nearest_neighbors_" + mode)(p.r_ids[0])
        deg, flat = _adj(p.edge_dicts[0], len(qs))
        ref_deg, ref_flat = g[f"fk_{mode}_deg"], g[f"fk_{mode}_adj"]
        # the only admissible differences come from float64 ties of the FK metric (< 1e-6)
        same = (deg == ref_deg).all() and (flat == ref_flat).all()
        if not same:
            off_a, off_b = np.concatenate([[0], np.cumsum(deg)]), np.concatenate([[0], np.cumsum(ref_deg)])
            diff = sum(list(flat[off_a[i]:off_a[i + 1]]) != list(ref_flat[off_b[i]:off_b[i + 1]]) for i in range(len(qs)))
            assert diff <= 2, f"{mode}: {diff} adjacency lists differ"


def test_load_roadmap_766(cuda, golden, tmp_path):
    """utils/pb_data_utils.load_roadmap + DecoupledPRM.load_roadmaps on the reference's prebuilt
    766-node Panda roadmap (golden = the reference's own parse of the CSV)."""
    from mmmp_b200.planners.prm.pybullet.decoupled.cbsprm import CBSPRM
    from mmmp_b200.utils.roadmap_io import load_roadmap, save_roadmap
    g = golden("roadmap_766")
    q, deg, adj = g["q"], g["deg"], g["adj"]
    off = np.concatenate([[0], np.cumsum(deg)])
    rm = {tuple(q[i]): [tuple(q[j]) for j in adj[off[i]:off[i + 1]]] for i in range(len(q))}
    path = os.path.join(tmp_path, "roadmap.csv")
    save_roadmap(rm, path)
    again = load_roadmap(path)
    assert list(again.keys()) == list(rm.keys()) and all(again[k] == rm[k] for k in rm)
    env = _panda_env(2)
    prm = CBSPRM(env, load_roadmap=path, maxdist=3.0, k1=55, k2=50, build_type='kdtree', prm_type='distance',
                 n=1, t=1, time_step=0.1, local_step=0.02)
    assert prm.compute_combined_nr_nodes() == 2 * 766
    assert sum(len(v) for v in prm.edge_dicts[0].values()) == 4936       # 766 nodes, 4936 directed edges
    d, f = _adj(prm.edge_dicts[1], 766)
    assert (d == deg).all() and (f == adj).all()


def test_cbs_prm_paths_are_collision_free_under_hull_oracle(cuda):
    """North-star check 4: build two Panda roadmaps on the GPU, plan with CBS-PRM on the host and
    re-validate every discretised configuration with the hull oracle (the restatement of the
    reference's pybullet checker) -- self, ground plane and robot-robot."""
    from mmmp_b200 import core
    from mmmp_b200.planners.prm.pybullet.decoupled.cbsprm import CBSPRM
    from oracle.hull import PandaHullChecker
    env = _panda_env(2)
    np.random.seed(5)     # the 10 + 10 start / goal poses come from numpy's global stream, as in the reference
    prm = CBSPRM(env, load_roadmap=None, maxdist=0.1, k1=12, k2=8, build_type='n', prm_type='degree', n=400, t=0,
                 time_step=0.1, local_step=0.02, seed=5)
    assert prm.compute_combined_nr_nodes() == 2 * (400 + 49)
    for r in range(2):
        ed = prm.edge_dicts[r]
        assert all(i in ed[j] for i in ed for j in ed[i])
    solution, l_t, q_t = prm.query()
    assert solution, "CBS-PRM found no solution"
    bases = [m.base_matrix for m in env.robot_models]
    hull = PandaHullChecker(bases, [core.plane((0, 0, 1), 0.0)])
    paths = {r: prm.discretize_path_in_time(r, solution[r]) for r in solution}
    for r_index, r_id in enumerate(prm.r_ids):
        q = np.array([np.asarray(v, float) for v in paths[r_id].values()])
        assert not hull.config_collision(q, r_index).any()
    # pairwise, at the reference's sweep times (cbsprm.py:165-183)
    t_max = max(max(p.keys()) for p in paths.values())
    qa, qb = [], []
    for t in np.round(np.arange(0, t_max, prm.time_step), 1):
        pa, pb = paths[prm.r_ids[0]], paths[prm.r_ids[1]]
        qa.append(pa[min(pa.keys(), key=lambda x: abs(x - t))])
        qb.append(pb[min(pb.keys(), key=lambda x: abs(x - t))])
    assert not hull.robot_robot_collision(np.array(qa, float), np.array(qb, float), 0, 1).any()


def test_coupled_panda_knn_prm(cuda):
    from mmmp_b200.planners.prm.pybullet.coupled.degree_coupled_prm import KDTreeKNNCoupledPRM
    from oracle.spheres import SphereChecker
    from mmmp_b200 import core
    env = _panda_env(2)
    prm = KDTreeKNNCoupledPRM(env, max_edge_len=3.0, n_knn=6, n_nodes=300, seed=2)
    assert len(prm.nodes) == 300 and len(prm.c_space) == 14
    q = np.array([n.q for n in prm.nodes])
    ck = SphereChecker([m.model for m in env.robot_models], [m.base_matrix for m in env.robot_models],
                       [core.plane((0, 0, 1), 0.0)])
    assert not ck.composite_collision(q).any()
    # candidates are the float64 k nearest in the 14-D composite space
    D = np.sqrt(((q[:, None] - q[None]) ** 2).sum(-1))
    np.fill_diagonal(D, np.inf)
    assert np.abs(np.sort(D, 1)[:, :6] - prm.last_candidates["dist"]).max() < 1e-12
    assert all(i in prm.roadmap[j] for i in prm.roadmap for j in prm.roadmap[i])


# ------------------------------------------------------------------ connected components
def _host_components(n, edges):
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for a, b in edges:
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    return np.array([find(i) for i in range(n)])


def test_connected_components_kernel_vs_union_find(cuda, golden):
    from mmmp_b200 import core
    rng = np.random.default_rng(5)
    for n, cap, p_edge in ((1, 3, 0.5), (257, 2, 0.3), (5000, 3, 0.25), (200000, 4, 0.12)):
        adj = rng.integers(0, n, size=(n, cap)).astype(np.int32)
        adj[rng.random((n, cap)) < 0.2] = -1                       # padding
        mask = (rng.random((n, cap)) < p_edge).astype(np.uint8)
        labels, count = core.connected_components(torch.from_numpy(adj).to(cuda), torch.from_numpy(mask).to(cuda))
        ii, cc = np.nonzero((adj >= 0) & (mask > 0))
        want = _host_components(n, zip(ii.tolist(), adj[ii, cc].tolist()))
        got = labels.cpu().numpy()
        assert (got == want).all()                                  # label = smallest node id of the component
        assert count == len(np.unique(want))
        labels2, count2 = core.connected_components(torch.from_numpy(np.where(mask > 0, adj, -1).astype(np.int32)).to(cuda))
        assert (labels2.cpu().numpy() == want).all() and count2 == count      # mask folded into the padding
    # the reference's prebuilt 766-node roadmap: components of its adjacency lists
    g = golden("roadmap_766")
    deg, flat = g["deg"], g["adj"]
    off = np.concatenate([[0], np.cumsum(deg)])
    adj = np.full((766, int(deg.max())), -1, dtype=np.int32)
    for i in range(766):
        adj[i, :deg[i]] = flat[off[i]:off[i + 1]]
    labels, count = core.connected_components(torch.from_numpy(adj).to(cuda))
    want = _host_components(766, [(i, int(j)) for i in range(766) for j in flat[off[i]:off[i + 1]]])
    assert (labels.cpu().numpy() == want).all() and count == len(np.unique(want))


def test_component_stitching_equals_the_sequential_loop(cuda):
    """compute_connected_components (GPU labels for large roadmaps) and the batched
    connect_connected_components_random reproduce the reference's sequential loops
    (decoupled_prm.py:613-704): same components, same added edges, same generator state."""
    import random
    from mmmp_b200.planners.prm.pybullet.decoupled.decoupled_prm import DecoupledPRM
    env = _panda_env(1)
    prm = DecoupledPRM(env, maxdist=0, k1=2, k2=1, build_type='n', prm_type='degree', n=1, t=1, time_step=0.1,
                       local_step=0.05)
    r_id = prm.r_ids[0]
    prm.add_nodes_robot(r_id, 400)
    prm.clear_edges(r_id)
    prm.connect_nearest_neighbors_degree(r_id)
    r_index = 0
    base_edges = {k: list(v) for k, v in prm.edge_dicts[r_index].items()}
    dfs = prm.compute_connected_components(r_id)
    prm.GPU_COMPONENTS_MIN_NODES = 1                                 # force the GPU path
    gpu = prm.compute_connected_components(r_id)
    assert [sorted(n.id for n in c) for c in gpu] == [sorted(n.id for n in c) for c in dfs]
    assert [c[0].id for c in gpu] == [c[0].id for c in dfs]         # components start at the same node
    assert len(dfs) > 3

    def sequential(components, n_pairs, max_tries):                  # the reference's loop, one check per try
        for i in range(len(components)):
            comp1 = components[i]
            if not comp1:
                continue
            for j in range(i + 1, len(components)):
                comp2 = components[j]
                if not comp2:
                    continue
                pairs, tries = [], 0
                if len(comp1) * len(comp2) > n_pairs:
                    while len(pairs) < n_pairs and tries < max_tries:
                        a, b = random.choice(comp1), random.choice(comp2)
                        if prm.is_collision_free_edge(a.q, b.q, r_id):
                            pairs.append((a, b))
                        tries += 1
                else:
                    for a in comp1:
                        for b in comp2:
                            if prm.is_collision_free_edge(a.q, b.q, r_id):
                                pairs.append((a, b))
                for a, b in pairs:
                    prm.edge_dicts[r_index][a.id].append(b.id)
                    prm.edge_dicts[r_index][b.id].append(a.id)
                if pairs:
                    comp1.extend(comp2)
                    components[j] = []
        return [c for c in components if c]

    random.seed(1234)
    seq = sequential([list(c) for c in dfs], 5, 20)
    seq_edges = {k: list(v) for k, v in prm.edge_dicts[r_index].items()}
    seq_state = random.getstate()
    prm.edge_dicts[r_index] = {k: list(v) for k, v in base_edges.items()}
    random.seed(1234)
    bat = prm.connect_connected_components_random([list(c) for c in dfs], r_id, n_pairs=5, max_tries=20)
    assert prm.edge_dicts[r_index] == seq_edges
    assert [[n.id for n in c] for c in bat] == [[n.id for n in c] for c in seq]
    assert random.getstate() == seq_state


# ------------------------------------------------------------------ task-space sampler
def test_task_space_sampler(cuda, golden):
    """sample_task_space_position draws exactly what the reference draws (golden = the reference's
    own function); sample_valid_in_task_space returns collision-free, 2-dp configurations whose
    tool sits on the sampled positions (IK residual + rounding), pointing down, and leaves numpy's
    global stream where the reference's one-pose-at-a-time loop would have left it."""
    from mmmp_b200.planners.prm.pybullet.decoupled.decoupled_prm import DecoupledPRM
    from oracle import ik
    g = golden("task_space_sampler")
    env = _panda_env(1)
    prm = DecoupledPRM(env, maxdist=0, k1=4, k2=3, build_type='n', prm_type='degree', n=1, t=1, time_step=0.1,
                       local_step=0.05)
    for c in range(3):
        np.random.seed(100 + c)
        got = np.array([prm.sample_task_space_position(g[f"start{c}"], g[f"goal{c}"]) for _ in range(200)])
        assert np.abs(got - g[f"pos{c}"]).max() < 1e-12
    r_id = prm.r_ids[0]
    model = prm.robot_models[0]
    np.random.seed(7)
    samples = np.array(prm.sample_valid_in_task_space(60, r_id, batch=32))
    state_after = np.random.get_state()
    assert samples.shape == (60, 7) and (samples == np.round(samples, 2)).all()
    assert not prm._configs_hit(0, samples, prm._flags()).any()
    # replay the reference's loop on the same stream: one position per try, accepted in order
    np.random.seed(7)
    start, goal = np.asarray(env.agents[0]['start'])[:7], np.asarray(env.agents[0]['goal'])[:7]
    sp, gp = model.position_from_fk(start), model.position_from_fk(goal)
    base = core_base(model)
    accepted, tries = [], 0
    while len(accepted) < 60:
        pos = prm.sample_task_space_position(sp, gp)
        tries += 1
        q, ok, _ = ik.ik_dls(model.model, base, pos[None], np.array([[1.0, 0, 0, 0]]), start[None])
        q = np.round(q[0], 2)
        if not prm._configs_hit(0, q, prm._flags())[0]:
            accepted.append((pos, q, bool(ok[0])))
    assert all(a == b for a, b in zip(np.random.get_state()[1].tolist(), state_after[1].tolist()))
    # the GPU sampler accepted the same tries; joint values agree with the float64 restatement
    ref_q = np.array([q for _, q, _ in accepted])
    conv = np.array([o for _, _, o in accepted])
    assert conv.mean() > 0.8
    assert np.abs(samples - ref_q)[conv].max() <= 0.011            # 2-dp rounding of fp32 vs fp64 iterates
    tool = np.array([model.position_from_fk(q) for q in samples])
    want = np.array([p for p, _, _ in accepted])
    assert np.linalg.norm(tool - want, axis=1)[conv].max() < 0.02  # 0.005 rad rounding x ~1 m reach
    down = np.array([model.orientation_from_fk(q)[:, 2] for q in samples])
    assert (down[conv, 2] < -0.99).all()                           # tool z axis points down (Euler 180, 0, 0)
    # the planner builds from this sampler when asked to
    prm2 = DecoupledPRM(env, maxdist=0, k1=4, k2=3, build_type='n', prm_type='degree', n=40, t=1, time_step=0.1,
                        local_step=0.05, sampler='task')
    assert len(prm2.node_lists[0]) >= 40


def core_base(model):
    from mmmp_b200 import core
    return core.base_matrix(model.base_position, model.base_orientation)


def test_start_goal_poses_are_sampled_in_task_space(cuda):
    """sample_start_poses / sample_goal_poses (decoupled_prm.py:438-470): tool positions drawn
    around the start / goal TOOL position (three normals per sample from numpy's global stream,
    sigma 0.2 m), closest IK from the start / goal pose with the tool pointing down.  The draws are
    the reference's; the joint values are pinned against the float64 restatement of the solver."""
    from mmmp_b200.planners.prm.pybullet.decoupled.decoupled_prm import DecoupledPRM
    from oracle import ik
    env = _panda_env(1)
    prm = DecoupledPRM(env, maxdist=0, k1=4, k2=3, build_type='n', prm_type='degree', n=1, t=1, time_step=0.1,
                       local_step=0.05)
    r_id, model = prm.r_ids[0], prm.robot_models[0]
    base = core_base(model)
    for which, fn in (("start", prm.sample_start_poses), ("goal", prm.sample_goal_poses)):
        pose = np.asarray(env.agents[0][which], dtype=float)[:7]
        np.random.seed(11)
        got = np.array(fn(r_id, n=10))
        after = np.random.get_state()[1].tolist()
        np.random.seed(11)
        pos0 = model.position_from_fk(pose)
        want = np.array([[pos0[0] + np.random.normal(0, 0.2), pos0[1] + np.random.normal(0, 0.2),
                          pos0[2] + np.random.normal(0, 0.2)] for _ in range(10)])
        assert np.random.get_state()[1].tolist() == after            # same number of draws, same order
        q_ref, ok, _ = ik.ik_dls(model.model, base, want, np.tile([[1.0, 0, 0, 0]], (10, 1)), np.tile(pose, (10, 1)))
        assert got.shape == (10, 7) and ok.sum() >= 5
        assert np.abs(got - q_ref)[ok].max() < 2e-3
        tool = np.array([model.position_from_fk(q) for q in got])
        assert np.linalg.norm(tool - want, axis=1)[ok].max() < 1e-3
        lo = np.array([iv.lower for iv in prm.c_spaces[0]])
        hi = np.array([iv.upper for iv in prm.c_spaces[0]])
        assert ((got >= lo - 1e-6) & (got <= hi + 1e-6)).all()


def test_prioritized_prm_paths_are_collision_free_under_hull_oracle(cuda):
    """The Prioritized-PRM drop-in class (prioritized_prm.py:14-49,113-168): two Panda roadmaps built
    on the GPU, robots planned in priority order on the host, the per-edge robot-robot sweep of
    robot_robot_collision_on_edge answered by the batched pair kernel.  Every discretised
    configuration of the returned paths is re-validated with the hull oracle (self, ground plane,
    robot-robot at the planner's own time grid), and the batched sweep is compared with the hull
    oracle's verdicts edge by edge (conservative: it may only ADD collisions)."""
    from mmmp_b200 import core
    from mmmp_b200.planners.prm.pybullet.decoupled.prioritized_prm import PrioritizedPRM
    from oracle.hull import PandaHullChecker
    env = _panda_env(2)
    np.random.seed(5)     # the 10 + 10 start / goal poses come from numpy's global stream, as in the reference
    prm = PrioritizedPRM(env, load_roadmap=None, maxdist=0.1, k1=12, k2=8, build_type='n', prm_type='degree', n=400, t=0,
                         time_step=0.1, local_step=0.02, seed=5)
    assert prm.compute_combined_nr_nodes() == 2 * (400 + 49)
    res = prm.query()
    assert res, "Prioritized-PRM found no solution"
    paths, l_t, q_t = res
    assert set(paths) == set(prm.r_ids) and l_t >= 0 and q_t >= 0
    for r_id, path in paths.items():
        assert path[0] == -1 and path[-1] == -2                         # start and goal node ids (decoupled_prm.py:943-995)
        ed = prm.edge_dicts[prm.r_ids.index(r_id)]
        assert all(b in ed[a] for a, b in zip(path[:-1], path[1:]))      # a walk on the roadmap
    # default priorities [1, 2]: np.argsort(...)[::-1] plans the LAST robot first (prioritized_prm.py:55-60)
    bases = [m.base_matrix for m in env.robot_models]
    hull = PandaHullChecker(bases, [core.plane((0, 0, 1), 0.0)])
    st = {r: prm.discretize_path_in_time(r, paths[r]) for r in paths}
    for r_index, r_id in enumerate(prm.r_ids):
        q = np.array([np.asarray(v, float) for v in st[r_id].values()])
        assert not hull.config_collision(q, r_index).any()
    t_max = max(max(p.keys()) for p in st.values())
    qa, qb = [], []
    for t in np.round(np.arange(0, t_max, prm.time_step), 1):
        pa, pb = st[prm.r_ids[0]], st[prm.r_ids[1]]
        qa.append(pa[min(pa.keys(), key=lambda x: abs(x - t))])
        qb.append(pb[min(pb.keys(), key=lambda x: abs(x - t))])
    assert not hull.robot_robot_collision(np.array(qa, float), np.array(qb, float), 0, 1).any()
    # robot_robot_collision_on_edge against the hull oracle on roadmap edges of robot 0 while robot 1
    # follows its path: never misses a hull collision
    r0, r1 = prm.r_ids
    ed = prm.edge_dicts[0]
    rng = np.random.default_rng(3)
    nodes = [n for n in ed if n >= 0 and ed[n]]
    missed = checked = 0
    for a in rng.choice(nodes, 40, replace=False):
        b = ed[a][0]
        start_time = float(rng.uniform(0, max(t_max - 1.0, 0.1)))
        got = prm.robot_robot_collision_on_edge(a, b, r0, {r1: st[r1]}, start_time)
        qa_, qb_ = np.array(prm.node_id_q_dicts[0][a]), np.array(prm.node_id_q_dicts[0][b])
        dt = np.linalg.norm(qa_ - qb_) / prm.local_step * prm.time_step
        t0 = np.round(np.ceil(start_time / prm.time_step) * prm.time_step, 1)
        t1 = np.round(np.floor((start_time + dt) / prm.time_step) * prm.time_step, 1)
        times = [np.round(t, 1) for t in np.arange(t0, t1 + prm.time_step, prm.time_step)]
        if not times:
            assert got is False
            continue
        mine = np.array([prm.find_q_in_q_t_interval_given_t(np.append(qa_, start_time), np.append(qb_, start_time + dt), t)
                         for t in times], float)[:, :7]
        keys = np.array(list(st[r1].keys()))
        vals = list(st[r1].values())
        theirs = np.array([np.asarray(vals[j], float) for j in np.abs(keys[None] - np.array(times)[:, None]).argmin(1)])
        ref = bool(hull.robot_robot_collision(mine, theirs, 0, 1).any())
        checked += 1
        missed += int(ref and not got)
    assert checked >= 10 and missed == 0


def test_kuka_collision_vs_sphere_oracle(cuda):
    """KUKA iiwa14 (the chain north_star names beside the Panda): configuration and edge collision of
    the reference's own sphere URDF (models/kuka_iiwa/urdf/iiwa14_spheres_collision.urdf) against the
    float64 sphere oracle -- self, obstacles, and two KUKAs against each other."""
    from mmmp_b200 import core
    from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF
    from oracle.spheres import SphereChecker
    m = core.load_model("kuka_iiwa14_spheres")
    lim = np.array(m["joint_limits"])
    rng = np.random.default_rng(8)
    bases = [core.base_matrix((0, 0, 0.0)), core.base_matrix((0.6, 0.1, 0.0), (0, 0, 1, 0))]
    obs = [core.plane((0, 0, 1), -0.2), core.box((0.45, 0.0, 0.4), (0.08, 0.3, 0.1)), core.sphere((-0.3, 0.3, 0.7), 0.15)]
    w1 = core.SpatialWorld([core.RobotSpec(m, bases[0])], obs)
    ck1 = SphereChecker([m], [bases[0]], obs)
    q = np.round(rng.uniform(lim[:, 0], lim[:, 1], (4000, 7)), 2)
    q32 = core.pack_f32(core.to_device_f64(q))
    for flags, ref in ((CHECK_SELF, ck1.config_collision(q, 0, obstacles=False)),
                       (CHECK_OBSTACLES, ck1.config_collision(q, 0, self_=False)),
                       (CHECK_SELF | CHECK_OBSTACLES, ck1.config_collision(q, 0))):
        got = w1.collide_configs(q32, 0, flags).cpu().numpy().astype(bool)
        assert (got != ref).sum() <= 3, (flags, int((got != ref).sum()))
    # (the URDF's own spheres of links two apart overlap in every configuration: a self check of this
    #  model always fires -- geometry/kuka_iiwa14_spheres.json self_pairs_note; obstacles decide here)
    full = ck1.config_collision(q, 0, self_=False)
    assert 0.02 < full.mean() < 0.98
    free = np.nonzero(~full)[0]
    ea, eb = rng.choice(free, 300), rng.choice(free, 300)
    edges = torch.tensor(np.stack([ea, eb], 1), dtype=torch.int32, device=cuda)
    got = w1.collide_edges(q32, edges, 20, 0.05, 0, CHECK_OBSTACLES).cpu().numpy().astype(bool)
    ref = ck1.edge_collision(q[ea], q[eb], 20, 0.05, 0, self_=False)
    assert (got != ref).sum() <= 3 and 0.05 < ref.mean() < 0.95
    w2 = core.SpatialWorld([core.RobotSpec(m, bases[0], 0), core.RobotSpec(m, bases[1], 7)], obs)
    ck2 = SphereChecker([m, m], bases, obs)
    qq = np.round(rng.uniform(np.tile(lim[:, 0], 2), np.tile(lim[:, 1], 2), (1500, 14)), 2)
    got = w2.collide_configs(core.pack_f32(core.to_device_f64(qq)), -1, CHECK_ROBOTS).cpu().numpy().astype(bool)
    ref = ck2.robot_robot_collision(qq[:, :7], qq[:, 7:], 0, 1)
    assert (got != ref).sum() <= 2 and 0.02 < ref.mean() < 0.98
