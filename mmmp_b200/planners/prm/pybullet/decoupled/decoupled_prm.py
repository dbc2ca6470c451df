"""Decoupled PRM for Panda / KUKA arms: the reference's DecoupledPRM interface
(planners/prm/pybullet/decoupled/decoupled_prm.py:25-1177) with the learning phase on the
GPU.  Same constructor, attributes (node_lists, edge_dicts, node_id_q_dicts,
node_q_id_dicts) and methods; the O(N^2) Python distance loops and the per-configuration
pybullet queries are replaced by batched kernels behind the C ABI:

  sampling                      -> mmmp_sample_uniform + mmmp_collide_configs + mmmp_compact_free
  k1 nearest candidates         -> mmmp_fk_features / mmmp_knn / mmmp_knn_refine
  is_collision_free_edge (all)  -> mmmp_collide_edges
  greedy accept pass            -> mmmp_greedy_degree_connect (native host replay)

Deviations from the reference, all outside the hot path: by default samples come from the
joint-space sampler (reference :310-321); `sampler='task'` selects the reference's task-space
sampler (:323-344) with the batched damped-least-squares IK kernel (mmmp_ik_dls) in place of
pybullet's IK (same sampled tool positions, joint values not pinned against pybullet --
SURVEY.md 8f rank 4); start/goal neighbourhood samples are joint-space perturbations; the
joint-space random stream is the library's Philox generator seeded by the extra `seed`
argument.
"""
from __future__ import annotations

import random
import time

import numpy as np
import torch

from ..... import core
from ....._lib import CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF
from .....robots.panda import Panda
from .....utils.roadmap_io import load_roadmap
from ..utils import Node

PI = np.pi
# the reference's 29 hand-picked "safe" poses (decoupled_prm.py:371-437), as multiples of pi
_SAFE = [
    (0, -1/4, 0, -3/4, 0, 1/2), (1/12, -1/4, 0, -3/4, 0, 1/2), (1/6, -1/4, 0, -3/4, 0, 1/2),
    (1/4, -1/4, 0, -3/4, 0, 1/2), (-1/12, -1/4, 0, -3/4, 0, 1/2), (-1/6, -1/4, 0, -3/4, 0, 1/2),
    (-1/4, -1/4, 0, -3/4, 0, 1/2), (0, -1/4, 0, -1/2, 0, 1/4), (0, -1/6, 0, -1/2, 0, 1/3),
    (0, -1/6, 0, -2/3, 0, 1/2), (0, 0, 0, -1/6, 0, 1/6), (0, 0, 0, -1/12, 0, 1/12),
]
_SAFE_LOW = [(0, -1/4, 0, -15/18), (0, -1/6, 0, -15/18), (0, -1/8, 0, -15/18), (0, -1/12, 0, -15/18),
             (0, 0, 0, -15/18), (1/6, -1/3, 0, -2/3), (1/8, -1/3, 0, -5/6), (-1/6, -1/3, 0, -2/3),
             (-1/8, -1/3, 0, -5/6), (0, -1/4, 1/6, -15/18), (0, -1/6, 1/6, -15/18), (1/2, -1/4, 0, -3/4),
             (1/3, -1/4, 0, -3/4), (-1/2, -1/4, 0, -3/4), (-1/3, -1/4, 0, -3/4), (0, -1/4, 0, -5/6),
             (0, -1/4, 0, -2/3)]


def safe_poses():
    """29 poses: q7 = pi/4 everywhere; the second family sets q6 = q2 - q4."""
    out = [[a * PI, b * PI, c * PI, d * PI, e * PI, f * PI, PI / 4] for a, b, c, d, e, f in _SAFE]
    for a, b, c, d in _SAFE_LOW:
        out.append([a * PI, b * PI, c * PI, d * PI, 0, (b * PI) - (d * PI), PI / 4])
    return out


class DecoupledPRM:
    def __init__(self, environment, load_roadmap=None, maxdist=0, k1=0, k2=0, build_type='n', prm_type='degree',
                 n=0, t=0, time_step=0., local_step=0., seed=0, sampler='joint') -> None:
        if sampler not in ('joint', 'task'):
            raise ValueError("sampler must be 'joint' (decoupled_prm.py:310-321) or 'task' (:323-344)")
        self.sampler = sampler
        self.env = environment
        self.agents = environment.agents
        self.robot_models = environment.robot_models
        self.obstacles = environment.obstacles
        self.r_ids = [m.r_id for m in environment.robot_models]
        self.c_spaces = self.create_c_spaces(environment.robot_models)

        self.maxdist = self.validate_positive_float(maxdist, 'maxdist') if prm_type == 'distance' else maxdist
        self.k1 = self.validate_positive_integer(k1, 'k1') if prm_type == 'degree' else k1
        self.k2 = self.validate_positive_integer(k2, 'k2') if prm_type == 'degree' else k2
        self.build_type = self.validate_build_type(build_type)
        self.prm_type = self.validate_prm_type(prm_type)
        self.n = self.validate_positive_integer(n, 'n') if build_type in ['n', 'kdtree'] else n
        self.t = self.validate_positive_float(t, 't') if build_type == 't' else t
        self.time_step = self.validate_positive_float(time_step, 'time_step')
        self.local_step = self.validate_positive_float(local_step, 'local_step')
        self.seed = int(seed)
        self._draws = 0          # Philox counter offset: every draw advances it
        self._rng = np.random.default_rng(self.seed)

        self.node_lists = [[] for _ in self.robot_models]
        self.edge_dicts = [{} for _ in self.robot_models]
        self.node_id_q_dicts = [{} for _ in self.robot_models]
        self.node_q_id_dicts = [{} for _ in self.robot_models]

        if load_roadmap is not None:
            self.load_roadmaps(load_roadmap)
        else:
            self.generate_roadmaps()

    # ------------------------------------------------------------------ validation (reference :130-159)
    def create_c_spaces(self, robot_models):
        return [m.arm_c_space if isinstance(m, Panda) else m.c_space for m in robot_models]

    @staticmethod
    def validate_positive_integer(value, name):
        assert isinstance(value, int) and value > 0, f"{name} must be a positive integer."
        return value

    @staticmethod
    def validate_positive_float(value, name):
        assert isinstance(value, float) and value > 0, f"{name} must be a positive float."
        return value

    @staticmethod
    def validate_build_type(build_type):
        assert build_type in ['t', 'n', 'kdtree'], "build_type must be 't' or 'n' or 'kdtree'."
        return build_type

    @staticmethod
    def validate_prm_type(prm_type):
        assert prm_type in ['degree', 'distance'], "prm_type must be 'degree' or 'distance'."
        return prm_type

    def generate_node_id_q_dict(self, nodes):
        return {node.id: node.q for node in nodes}

    def generate_node_q_id_dict(self, nodes):
        return {tuple(node.q): node.id for node in nodes}

    # ------------------------------------------------------------------ device helpers
    @property
    def n_steps(self) -> int:
        return len(np.arange(0, 1, self.local_step))

    def _flags(self):
        return CHECK_SELF | CHECK_OBSTACLES      # is_collision_free_sample: self + obstacles (:840-850)

    def _device_nodes(self, r_index):
        q = np.array([node.q for node in self.node_lists[r_index]], dtype=np.float64).reshape(-1, len(self.c_spaces[r_index]))
        q64 = core.to_device_f64(q)
        return q64, core.pack_f32(q64)

    def _features(self, r_index, q64, q32):
        """Row sets of the planner's metric: specific_distance -> model.distance_metric (:827-830)."""
        return self.env.world.fk_features(q64, q32, r_index)

    def _append_nodes(self, r_index, samples):
        start = len(self.node_lists[r_index])
        nodes = [Node(start + i, s) for i, s in enumerate(samples)]
        self.node_lists[r_index].extend(nodes)
        self.edge_dicts[r_index].update({node.id: [] for node in nodes})
        self.node_id_q_dicts[r_index].update(self.generate_node_id_q_dict(nodes))
        self.node_q_id_dicts[r_index].update(self.generate_node_q_id_dict(nodes))
        return nodes

    # ------------------------------------------------------------------ learning
    def load_roadmaps(self, roadmap_file):
        """:208-224 -- same file for every robot, q re-rounded to 2 dp, no collision checks."""
        for r_index in range(len(self.r_ids)):
            roadmap = load_roadmap(roadmap_file)
            for i, q in enumerate(roadmap.keys()):
                node = Node(i, np.round(np.array(q), 2))
                self.node_lists[r_index].append(node)
                self.edge_dicts[r_index].update({node.id: []})
            self.node_id_q_dicts[r_index] = self.generate_node_id_q_dict(self.node_lists[r_index])
            self.node_q_id_dicts[r_index] = self.generate_node_q_id_dict(self.node_lists[r_index])
            for i, q in enumerate(roadmap.keys()):
                for nb in roadmap[q]:
                    nb = np.round(np.array(nb), 2)
                    self.edge_dicts[r_index][i].append(self.node_q_id_dicts[r_index][tuple(nb)])

    def generate_roadmaps(self):
        for r_id in self.r_ids:
            self.generate_roadmap(r_id)

    def generate_roadmap(self, r_id):
        r_index = self.r_ids.index(r_id)
        if self.build_type == 't':
            raise NotImplementedError("Update roadmap based on time is not implemented.")
        samples = (self.sample_valid_in_task_space(self.n, r_id) if self.sampler == 'task'
                   else self.sample_valid_in_joint_space_random(self.n, r_id))
        self._append_nodes(r_index, samples)
        if self.prm_type == 'degree':
            self.connect_nearest_neighbors_degree(r_id)
        else:
            self.connect_nearest_neighbors_distance(r_id)
        for extra in (self.sample_start_poses(r_id, n=10), self.sample_goal_poses(r_id, n=10),
                      self.sample_safe_poses(r_id)):
            nodes = self._append_nodes(r_index, extra)
            if self.prm_type == 'degree':
                self.connect_extra_nodes_degree(r_id, nodes)
            else:
                self.connect_extra_nodes_distance(r_id, nodes)
        components = self.compute_connected_components(r_id)
        self.connect_connected_components_random(components, r_id, n_pairs=5, max_tries=50)

    # ------------------------------------------------------------------ sampling
    def sample_valid_in_joint_space_random(self, n, r_id):
        """uniform in the joint limits, rounded to 2 dp, rejected on collision (:310-321); batched."""
        r_index = self.r_ids.index(r_id)
        lo = np.array([iv.lower for iv in self.c_spaces[r_index]])
        hi = np.array([iv.upper for iv in self.c_spaces[r_index]])
        got, have, rate = [], 0, 0.5
        while have < n:
            m = int(np.ceil((n - have) / rate * 1.1)) + 32
            q64, q32 = core.sample_uniform(lo, hi, m, self.seed + 7919 * r_index, self._draws)
            self._draws += m
            hit = self.env.world.collide_configs(q32, r_index, self._flags())
            free = core.gather_rows(q64, core.compact_free(hit))
            rate = max(free.shape[0] / m, 1e-3)
            got.append(free)
            have += free.shape[0]
        q = torch.cat(got, 0)[:n].cpu().numpy()
        return [row for row in q]


    def sample_task_space_position(self, start_pos, goal_pos, k=1, mu=0.4, sigma=0.2):
        """:346-369, same draws from numpy's global stream in the same order (t, chi-square, theta)."""
        start_pos, goal_pos = np.array(start_pos, dtype=float), np.array(goal_pos, dtype=float)
        W = goal_pos - start_pos
        W /= np.linalg.norm(W)
        U = np.array([0., 1., 0.]) if np.allclose(W, [1., 0., 0.]) else np.array([1., 0., 0.])
        U -= np.dot(U, W) * W
        U /= np.linalg.norm(U)
        V = np.cross(W, U)
        t = np.random.uniform(0, 1)
        d = mu / k * (sigma ** 2 / (2 * k)) ** (1 / 2) * np.random.chisquare(k)
        theta = np.random.uniform(0, 2 * np.pi)
        return start_pos + t * (goal_pos - start_pos) + d * (np.cos(theta) * U + np.sin(theta) * V)

    def sample_valid_in_task_space(self, n, r_id, batch=None):
        """:323-344: positions in a tube around the start-goal segment of the tool, tool pointing
        down (quaternion of Euler (180 deg, 0, 0)), closest IK from the start pose, rounded to 2 dp,
        kept when collision free.  The reference solves one pose per pybullet call; here a batch
        of positions is drawn ahead (same global numpy stream, rewound to what the one-at-a-time
        loop would have consumed), solved by ONE mmmp_ik_dls launch and checked by one
        collision launch.  pybullet's IK arithmetic is not available: joint values differ from
        the reference's (parity unpinned), the sampled tool positions do not."""
        r_index = self.r_ids.index(r_id)
        model = self.robot_models[r_index]
        start = np.asarray(self.agents[r_index]['start'], dtype=float)[:7]
        goal = np.asarray(self.agents[r_index]['goal'], dtype=float)[:7]
        start_pos, goal_pos = model.position_from_fk(start), model.position_from_fk(goal)
        quat = (1.0, 0.0, 0.0, 0.0)                               # p.getQuaternionFromEuler([pi, 0, 0])
        dev = core._require_cuda()
        q0 = torch.tensor(start[None], dtype=torch.float64, device=dev)
        samples = []
        while len(samples) < n:
            m = int(batch or max(64, 2 * (n - len(samples))))
            state = np.random.get_state()
            pos = np.array([self.sample_task_space_position(start_pos, goal_pos) for _ in range(m)])
            tq = torch.tensor(np.tile(quat, (m, 1)), dtype=torch.float64, device=dev)
            q, _, _ = self.env.world.ik(torch.tensor(pos, dtype=torch.float64, device=dev), tq, q0, r_index)
            q = np.round(q.cpu().numpy(), 2)
            free = ~self._configs_hit(r_index, q, self._flags())
            used = 0
            for qi, ok in zip(q, free):
                if len(samples) >= n:
                    break
                used += 1
                if ok:
                    samples.append(qi)
            np.random.set_state(state)
            for _ in range(used):
                self.sample_task_space_position(start_pos, goal_pos)
        return samples

    def sample_safe_poses(self, r_id, sigma=0.2):
        return safe_poses()

    def sample_around_task_space_position(self, pos, sigma):
        """:465-470: three draws from numpy's global stream, x then y then z."""
        x, y, z = pos
        x += np.random.normal(0, sigma)
        y += np.random.normal(0, sigma)
        z += np.random.normal(0, sigma)
        return [x, y, z]

    def _sample_around_pose(self, r_id, pose, n):
        """sample_start_poses / sample_goal_poses (:438-462): n tool positions drawn around the
        tool position of `pose` (sigma 0.2 m, same draws in the same order as the reference),
        each solved by closest IK from `pose` with the tool pointing down (quaternion of Euler
        (180 deg, 0, 0)) -- one batched mmmp_ik_dls launch instead of n pybullet calls.  The tool
        positions are the reference's; the joint values come from this library's DLS solver
        (pybullet's IK arithmetic is not available: parity unpinned)."""
        r_index = self.r_ids.index(r_id)
        model = self.robot_models[r_index]
        pose = np.asarray(pose, dtype=float)[:7]
        pos0 = model.position_from_fk(pose)
        targets = np.array([self.sample_around_task_space_position(pos0, 0.2) for _ in range(n)], dtype=np.float64)
        dev = core._require_cuda()
        tq = torch.tensor(np.tile((1.0, 0.0, 0.0, 0.0), (n, 1)), dtype=torch.float64, device=dev)
        q0 = torch.tensor(pose[None], dtype=torch.float64, device=dev)
        q, _, _ = self.env.world.ik(torch.tensor(targets, dtype=torch.float64, device=dev), tq, q0, r_index)
        return [row for row in q.cpu().numpy()]

    def sample_start_poses(self, r_id, n=10):
        return self._sample_around_pose(r_id, self.agents[self.r_ids.index(r_id)]['start'], n)

    def sample_goal_poses(self, r_id, n=10):
        return self._sample_around_pose(r_id, self.agents[self.r_ids.index(r_id)]['goal'], n)

    # ------------------------------------------------------------------ connection
    def _edges_hit(self, r_index, q32, edges):
        """[E] uint8: 1 where the straight edge (q32[a], q32[b]) is NOT collision free."""
        return self.env.world.collide_edges(q32, edges, self.n_steps, self.local_step, r_index, self._flags())

    def _candidate_verdicts(self, r_index, q32, idx):
        """[n, k] uint8 on the device: 1 where is_collision_free_edge(node_i, candidate) holds.  Mutual
        candidates are checked once, low id -> high id: the first of the accept loop's two checks."""
        return self.env.world.collide_knn_edges(q32, idx, self.n_steps, self.local_step, r_index, self._flags())

    def connect_nearest_neighbors_degree(self, r_id):
        """:480-504.  k1 candidates per node (all other nodes: the reference's degree filter
        compares a list with an int and never fires), every candidate edge checked on the
        GPU, then the sequential accept pass replayed natively."""
        r_index = self.r_ids.index(r_id)
        nodes = self.node_lists[r_index]
        if len(nodes) < 2:
            return
        q64, q32 = self._device_nodes(r_index)
        F = self._features(r_index, q64, q32)
        k1 = min(self.k1, len(nodes) - 1)
        idx, dist, amb = core.knn_exact(F, F, k1)
        free_d = self._candidate_verdicts(r_index, q32, idx)
        # the order-dependent accept pass, on the device (same adjacency and list order as the
        # sequential loop: tests/test_gpu_adjacency.py)
        adj, deg = core.greedy_degree_connect_device(idx, free_d.contiguous(), self.k2)
        adj, deg = adj.cpu().numpy(), deg.cpu().numpy()
        self.last_candidates = {"idx": idx.cpu().numpy(), "dist": dist.cpu().numpy(), "edge_free": free_d.cpu().numpy(),
                                "ambiguous": amb.cpu().numpy()}
        ed = self.edge_dicts[r_index]
        for i, node in enumerate(nodes):
            ed[node.id] = [nodes[j].id for j in adj[i, :deg[i]]]

    def connect_extra_nodes_degree(self, r_id, nodes):
        """:527-550 -- the new nodes look at ALL nodes, no degree cap on the other end."""
        r_index = self.r_ids.index(r_id)
        if not nodes:
            return
        all_nodes = self.node_lists[r_index]
        q64, q32 = self._device_nodes(r_index)
        F = self._features(r_index, q64, q32)
        first = nodes[0].id
        k1 = min(self.k1, len(all_nodes) - 1)
        idx, _, _ = core.knn_exact(F, F.rows(first, first + len(nodes)), k1, query_id0=first)
        edges = core.edges_from_knn(idx)
        edges[:, 0] += first
        hit = self._edges_hit(r_index, q32, edges)
        free = (hit == 0).reshape(idx.shape).cpu().numpy()
        idx = idx.cpu().numpy()
        ed = self.edge_dicts[r_index]
        for a, node in enumerate(nodes):
            for c in range(idx.shape[1]):
                if len(ed[node.id]) == self.k2:
                    break
                j = int(idx[a, c])
                if j < 0:
                    continue
                nb = all_nodes[j]
                if free[a, c] and nb.id not in ed[node.id]:
                    ed[node.id].append(nb.id)
                    ed[nb.id].append(node.id)

    def _distance_candidates(self, r_index, q64, q32, lo, hi):
        """rows lo..hi: neighbours with 0 < d < maxdist ascending (dist, id) + edge verdicts."""
        F = self._features(r_index, q64, q32)
        off, idx, dist = core.radius_search(F, F.rows(lo, hi), float(self.maxdist), strict=True, exclude_self=False,
                                            exclude_zero=True, query_id0=lo)
        counts = (off[1:] - off[:-1])
        src = torch.repeat_interleave(torch.arange(lo, hi, device=idx.device, dtype=torch.int32), counts)
        edges = torch.stack([src, idx], 1).contiguous()
        hit = self._edges_hit(r_index, q32, edges)
        off, idx, dist, free = off.cpu().numpy(), idx.cpu().numpy(), dist.cpu().numpy(), (hit == 0).cpu().numpy()
        rows = []
        for a in range(hi - lo):
            s, e = off[a], off[a + 1]
            order = np.lexsort((idx[s:e], dist[s:e]))
            rows.append((idx[s:e][order], free[s:e][order]))
        return rows

    def connect_nearest_neighbors_distance(self, r_id):
        """:553-571"""
        r_index = self.r_ids.index(r_id)
        nodes = self.node_lists[r_index]
        if len(nodes) < 2:
            return
        q64, q32 = self._device_nodes(r_index)
        rows = self._distance_candidates(r_index, q64, q32, 0, len(nodes))
        ed = self.edge_dicts[r_index]
        for node, (cand, free) in zip(nodes, rows):
            for j, ok in zip(cand, free):
                nb = nodes[int(j)]
                if ok and nb.id not in ed[node.id]:
                    ed[node.id].append(nb.id)
                    ed[nb.id].append(node.id)

    def connect_extra_nodes_distance(self, r_id, nodes):
        """:592-610"""
        r_index = self.r_ids.index(r_id)
        if not nodes:
            return
        all_nodes = self.node_lists[r_index]
        q64, q32 = self._device_nodes(r_index)
        first = nodes[0].id
        rows = self._distance_candidates(r_index, q64, q32, first, first + len(nodes))
        ed = self.edge_dicts[r_index]
        for node, (cand, free) in zip(nodes, rows):
            for j, ok in zip(cand, free):
                nb = all_nodes[int(j)]
                if ok and nb.id not in ed[node.id]:
                    ed[node.id].append(nb.id)
                    ed[nb.id].append(node.id)

    # ------------------------------------------------------------------ components (:613-704)
    # roadmaps above this size get their components from the GPU union-find (same partition, the
    # components in the reference's order, nodes inside a component by id instead of DFS order)
    GPU_COMPONENTS_MIN_NODES = 50000

    def component_labels(self, r_id):
        """labels[i] = position (in the node list) of the first node of node i's component, on the
        GPU (mmmp_connected_components); the partition of compute_connected_components (:613-636)."""
        r_index = self.r_ids.index(r_id)
        nodes = self.node_lists[r_index]
        pos = {n.id: i for i, n in enumerate(nodes)}
        edges = self.edge_dicts[r_index]
        cap = max(1, max((len(edges.get(n.id, ())) for n in nodes), default=1))
        adj = np.full((len(nodes), cap), -1, dtype=np.int32)
        for i, n in enumerate(nodes):
            row = [pos[j] for j in edges.get(n.id, ()) if j in pos]
            adj[i, :len(row)] = row
        labels, _ = core.connected_components(torch.from_numpy(adj).to(core._require_cuda()))
        return labels.cpu().numpy()

    def compute_connected_components(self, r_id):
        """Depth-first components in the reference's visiting order (stack, last in first out);
        large roadmaps: the same components from the GPU labels, nodes in list order."""
        r_index = self.r_ids.index(r_id)
        nodes = self.node_lists[r_index]
        if len(nodes) >= self.GPU_COMPONENTS_MIN_NODES:
            labels = self.component_labels(r_id)
            order = np.argsort(labels, kind="stable")            # by first node, then list order
            cuts = np.nonzero(np.diff(labels[order]))[0] + 1
            return [[nodes[i] for i in part] for part in np.split(order, cuts)]
        by_id = {n.id: n for n in nodes}
        visited, components = set(), []
        for node in nodes:
            if node.id in visited:
                continue
            component, stack = [], [node]
            while stack:
                current = stack.pop()
                if current.id in visited:
                    continue
                visited.add(current.id)
                component.append(current)
                for nid in self.edge_dicts[r_index][current.id]:
                    if nid in by_id and nid not in visited:
                        stack.append(by_id[nid])
            components.append(component)
        return components

    def connect_connected_components_random(self, components, r_id, n_pairs=5, max_tries=20):
        """:666-704.  The random tries of one component pair are drawn ahead (the draws do not depend
        on the verdicts), checked in ONE batched launch, and the generator is then rewound to where
        the reference's early-stopping loop would have left it: same edges, same later draws."""
        r_index = self.r_ids.index(r_id)
        for i in range(len(components)):
            comp1 = components[i]
            if not comp1:
                continue
            for j in range(i + 1, len(components)):
                comp2 = components[j]
                if not comp2:
                    continue
                if len(comp1) * len(comp2) > n_pairs:
                    state = random.getstate()
                    draws = [(random.choice(comp1), random.choice(comp2)) for _ in range(max_tries)]
                    free = self._segments_free(r_index, [a.q for a, _ in draws], [b.q for _, b in draws])
                    connecting_pairs, used = [], 0
                    for (a, b), ok in zip(draws, free):
                        if len(connecting_pairs) >= n_pairs:
                            break
                        used += 1
                        if ok:
                            connecting_pairs.append((a, b))
                    random.setstate(state)
                    for _ in range(used):
                        random.choice(comp1)
                        random.choice(comp2)
                else:
                    pairs = [(a, b) for a in comp1 for b in comp2]
                    free = self._segments_free(r_index, [a.q for a, _ in pairs], [b.q for _, b in pairs])
                    connecting_pairs = [p for p, ok in zip(pairs, free) if ok]
                for node1, node2 in connecting_pairs:
                    self.edge_dicts[r_index][node1.id].append(node2.id)
                    self.edge_dicts[r_index][node2.id].append(node1.id)
                if connecting_pairs:
                    comp1.extend(comp2)
                    components[j] = []
        return [comp for comp in components if comp]

    # ------------------------------------------------------------------ roadmap updates (:707-775)
    def clear_edges(self, r_id):
        r_index = self.r_ids.index(r_id)
        for node in self.node_lists[r_index]:
            self.edge_dicts[r_index].update({node.id: []})

    def update_roadmaps_edges(self):
        for r_id in self.r_ids:
            self.update_roadmap_edges(r_id)

    def update_roadmap_edges(self, r_id):
        self.clear_edges(r_id)
        if self.build_type == 't':
            raise NotImplementedError("Update roadmap based on time is not implemented.")
        if self.prm_type == 'degree':
            self.connect_nearest_neighbors_degree(r_id)
        else:
            self.connect_nearest_neighbors_distance(r_id)
        components = self.compute_connected_components(r_id)
        self.connect_connected_components_random(components, r_id, n_pairs=5, max_tries=20)

    def add_nodes(self, n):
        for r_id in self.r_ids:
            self.add_nodes_robot(r_id, n)

    def add_nodes_robot(self, r_id, n):
        self._append_nodes(self.r_ids.index(r_id), self.sample_valid_in_joint_space_random(n, r_id))

    def add_nodes_and_update_edges(self, n):
        for r_id in self.r_ids:
            self.add_nodes_robot(r_id, n)
            self.update_roadmap_edges(r_id)

    def remove_nodes_robot(self, r_id, n):
        r_index = self.r_ids.index(r_id)
        for _ in range(n):
            node = self.node_lists[r_index].pop(-1)
            self.node_id_q_dicts[r_index].pop(node.id)
            self.node_q_id_dicts[r_index].pop(node.q)

    def remove_nodes(self, n):
        for r_id in self.r_ids:
            self.remove_nodes_robot(r_id, n)

    def remove_nodes_and_update_edges(self, n):
        for r_id in self.r_ids:
            self.remove_nodes_robot(r_id, n)
            self.update_roadmap_edges(r_id)

    # ------------------------------------------------------------------ distance (:821-837)
    def distance(self, r_id, q1, q2):
        return self.specific_distance(r_id, q1, q2)

    def general_distance(self, q1, q2):
        return np.linalg.norm(np.array(q1) - np.array(q2))

    def specific_distance(self, r_id, q1, q2):
        return self.robot_models[self.r_ids.index(r_id)].distance_metric(q1, q2)

    def task_space_distance(self, r_id, q1, q2):
        model = self.robot_models[self.r_ids.index(r_id)]
        return np.linalg.norm(model.position_from_fk(q1) - model.position_from_fk(q2))

    # ------------------------------------------------------------------ collision (:840-876)
    def _configs_hit(self, r_index, qs, flags):
        q32 = core.pack_f32(core.to_device_f64(np.atleast_2d(np.asarray(qs, dtype=np.float64))))
        return self.env.world.collide_configs(q32, r_index, flags).cpu().numpy().astype(bool)

    def _segments_free(self, r_index, qa, qb):
        a = core.pack_f32(core.to_device_f64(np.atleast_2d(np.asarray(qa, dtype=np.float64))))
        b = core.pack_f32(core.to_device_f64(np.atleast_2d(np.asarray(qb, dtype=np.float64))))
        hit = self.env.world.collide_segments(a, b, self.n_steps, self.local_step, r_index, self._flags())
        return (hit == 0).cpu().numpy()

    def is_collision_free_sample(self, sample, r_id):
        r_index = self.r_ids.index(r_id)
        self.robot_models[r_index].set_arm_pose(sample)
        return not self._configs_hit(r_index, sample, self._flags())[0]

    def is_collision_free_node(self, node, r_id):
        return self.is_collision_free_sample(node.q, r_id)

    def is_collision_free_edge(self, q1, q2, r_id):
        return bool(self._segments_free(self.r_ids.index(r_id), q1, q2)[0])

    def robot_self_collision(self, q, r_id):
        r_index = self.r_ids.index(r_id)
        self.robot_models[r_index].set_arm_pose(q)
        return bool(self._configs_hit(r_index, q, CHECK_SELF)[0])

    def robot_robot_collision(self, q1, q2, r_1_id, r_2_id):
        a, b = self.r_ids.index(r_1_id), self.r_ids.index(r_2_id)
        self.robot_models[a].set_arm_pose(q1)
        self.robot_models[b].set_arm_pose(q2)
        return self.robot_robot_collision_batch([q1], [q2], a, b)[0] or None

    def robot_robot_collision_batch(self, qa, qb, a, b):
        """Batched form used by the query phase: [n] verdicts for pairs (qa[i], qb[i])."""
        ta = core.pack_f32(core.to_device_f64(np.atleast_2d(np.asarray(qa, dtype=np.float64))))
        tb = core.pack_f32(core.to_device_f64(np.atleast_2d(np.asarray(qb, dtype=np.float64))))
        return self.env.world.collide_robot_pairs(ta, tb, a, b).cpu().numpy().astype(bool)

    # ------------------------------------------------------------------ interpolation (:879-909)
    def find_q_in_q_t_interval_given_t(self, q1_t1, q2_t2, t):
        delta_t = q2_t2[-1] - q1_t1[-1]
        x = (t - q1_t1[-1]) / delta_t
        return np.round(q1_t1[:-1] + (q2_t2[:-1] - q1_t1[:-1]) * x, 2)

    def discretize_path_in_time(self, r_id, id_path):
        r_index = self.r_ids.index(r_id)
        discretized_path = {}
        interval_end_t = 0.0
        end_config = None
        for i in range(len(id_path) - 1):
            interval_start_t = interval_end_t
            start_config = np.array(self.node_id_q_dicts[r_index][id_path[i]])
            end_config = np.array(self.node_id_q_dicts[r_index][id_path[i + 1]])
            interval_end_t = interval_start_t + np.linalg.norm(end_config - start_config) / self.local_step * self.time_step
            first = np.round(np.ceil(interval_start_t / self.time_step) * self.time_step, 1)
            for t in np.arange(first, interval_end_t, self.time_step):
                t = np.round(t, 1)
                discretized_path[t] = self.find_q_in_q_t_interval_given_t(np.append(start_config, interval_start_t),
                                                                          np.append(end_config, interval_end_t), t)
        discretized_path[interval_end_t] = end_config
        return discretized_path

    # ------------------------------------------------------------------ query hooks
    def query(self):
        raise NotImplementedError("Query method not implemented.")

    def query_robot(self, r_id):
        raise NotImplementedError("Query method not implemented.")

    def a_star_search(self, r_id):
        raise NotImplementedError("A* search method not implemented.")

    def add_start_goal_nodes(self, r_id):
        """:943-995 -- start = id -1, goal = id -2, k1 nearest roadmap nodes each, up to k2 free edges."""
        r_index = self.r_ids.index(r_id)
        start_config = self.agents[r_index]['start']
        goal_config = self.agents[r_index]['goal']
        if not self.is_collision_free_sample(start_config, r_id):
            print(f"Start configuration for robot {r_id} is in collision.")
            return
        if not self.is_collision_free_sample(goal_config, r_id):
            print(f"Goal configuration for robot {r_id} is in collision.")
            return
        nodes = self.node_lists[r_index]
        q64, q32 = self._device_nodes(r_index)
        F = self._features(r_index, q64, q32)
        qq = core.to_device_f64(np.array([start_config, goal_config], dtype=np.float64))
        qq32 = core.pack_f32(qq)
        Fq = self._features(r_index, qq, qq32)
        k1 = min(self.k1, len(nodes))
        idx, _, _ = core.knn_exact(F, Fq, k1, exclude_self=False)
        idx = idx.cpu().numpy()
        ed = self.edge_dicts[r_index]
        for row, (nid, cfg) in enumerate(((-1, start_config), (-2, goal_config))):
            ed.update({nid: []})
            cand = [nodes[int(j)] for j in idx[row] if j >= 0]
            free = self._segments_free(r_index, [cfg] * len(cand), [c.q for c in cand]) if cand else []
            for nb, ok in zip(cand, free):
                if len(ed[nid]) == self.k2:
                    break
                if ok:
                    ed[nid].append(nb.id)
                    ed[nb.id].append(nid)
            self.node_id_q_dicts[r_index].update({nid: tuple(cfg)})
            self.node_q_id_dicts[r_index].update({tuple(cfg): nid})

    def delete_start_goal_nodes(self, r_id):
        r_index = self.r_ids.index(r_id)
        ed = self.edge_dicts[r_index]
        for nid in (-1, -2):
            ed.pop(nid, None)
        for key in list(ed.keys()):
            ed[key] = [x for x in ed[key] if x not in (-1, -2)]
        for nid in (-1, -2):
            cfg = self.node_id_q_dicts[r_index].pop(nid)
            self.node_q_id_dicts[r_index].pop(cfg)

    def heuristic(self, n_id1, n_id2, r_id):
        r_index = self.r_ids.index(r_id)
        return self.distance(r_id, np.array(self.node_id_q_dicts[r_index][n_id1]),
                             np.array(self.node_id_q_dicts[r_index][n_id2]))

    # ------------------------------------------------------------------ statistics (:1121-1177)
    def compute_combined_avg_degree(self):
        total_nb = sum(len(nb) for ed in self.edge_dicts for nb in ed.values())
        return total_nb / sum(len(nl) for nl in self.node_lists)

    def compute_combined_nr_nodes(self):
        return sum(len(nl) for nl in self.node_lists)

    def compute_nr_edges(self, r_id):
        return sum(len(nb) for nb in self.edge_dicts[self.r_ids.index(r_id)].values()) / 2

    def compute_combined_nr_edges(self):
        return sum(self.compute_nr_edges(r_id) for r_id in self.r_ids)

    def compute_total_path_length(self, id_paths):
        total_length = 0
        for r_id, path in id_paths.items():
            d = self.node_id_q_dicts[self.r_ids.index(r_id)]
            q_prev = d[path[0]]
            for nid in path[1:]:
                total_length += self.task_space_distance(r_id, q_prev, d[nid])
                q_prev = d[nid]
        return total_length
