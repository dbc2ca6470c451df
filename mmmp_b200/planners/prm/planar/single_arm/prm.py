"""Planar PRM base (reference planners/prm/planar/single_arm/prm.py:22-347): one arm,
`roadmap` = {node id: [neighbour ids]}.  Sampling, collision predicates and neighbour
search are batched GPU calls; graph search is host code."""
from __future__ import annotations

import heapq

import numpy as np
import torch

from ..... import core
from ....._lib import CHECK_BOUNDS, CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF
from ..utils import INF, Node


class PRM:
    ROBOT = 0                                                   # arm the rows refer to (-1 = composite)
    FLAGS = CHECK_BOUNDS | CHECK_SELF | CHECK_OBSTACLES         # is_collision_free_sample :104-116

    def __init__(self, environment, local_step, seed=0) -> None:
        self.env = environment
        self.agents = environment.agents
        self.robot_models = environment.robot_models
        self.config_space = self.robot_models[0].config_space
        self.obstacles = environment.obstacles
        self.local_step = local_step
        self.nodes = []
        self.node_id_config_dict = {}
        self.node_config_id_dict = {}
        self.roadmap = {}
        self.seed = int(seed)
        self._draws = 0

    def update_node_id_config_dict(self, nodes):
        for node in nodes:
            self.node_id_config_dict[node.id] = node.q

    def update_node_config_id_dict(self, nodes):
        for node in nodes:
            self.node_config_id_dict[node.q] = node.id

    # ------------------------------------------------------------------ device helpers
    @property
    def n_steps(self) -> int:
        return len(np.arange(0, 1, self.local_step))

    def _device_nodes(self):
        return core.to_device_f64(np.array([n.q for n in self.nodes], dtype=np.float64).reshape(-1, len(self.config_space)))

    def _register(self, samples):
        new = []
        for s in samples:
            node = Node(id=len(self.nodes), q=s)
            self.nodes.append(node)
            self.roadmap[node.id] = []
            self.node_config_id_dict[node.q] = node.id
            self.node_id_config_dict[node.id] = node.q
            new.append(node)
        return new

    def _l2(self, q64):
        return core.l2_features(q64)

    def _fk_sq(self, q64):
        """specific_distance rows: stacked link end points of every arm, sum of squared norms."""
        parts, off = [], 0
        for a, m in enumerate(self.robot_models if self.ROBOT < 0 else [self.robot_models[self.ROBOT]]):
            arm = a if self.ROBOT < 0 else self.ROBOT
            n = m.n_links
            parts.append(self.env.world.fk(q64[:, off:off + n].contiguous(), arm).reshape(q64.shape[0], -1))
            off += n
        pos = torch.cat(parts, 1).contiguous()
        return core.l2_features(pos, squared=True)

    def _edges_free(self, q64, edges):
        hit = self.env.world.collide_edges(q64, edges, self.n_steps, self.local_step, self.ROBOT, self.FLAGS)
        return (hit == 0).cpu().numpy()

    # ------------------------------------------------------------------ sampling (:52-73)
    def generate_free_samples_array(self, n, free_only=True):
        lo = np.array([iv.lower for iv in self.config_space])
        hi = np.array([iv.upper for iv in self.config_space])
        got, have, rate = [], 0, 0.5
        while have < n:
            m = n if not free_only else int(np.ceil((n - have) / rate * 1.1)) + 32
            q64, _ = core.sample_uniform(lo, hi, m, self.seed, self._draws)
            self._draws += m
            if free_only:
                hit = self.env.world.collide_configs(q64, self.ROBOT, self.FLAGS)
                q64 = core.gather_rows(q64, core.compact_free(hit))
                rate = max(q64.shape[0] / m, 1e-3)
            got.append(q64)
            have += q64.shape[0]
        return torch.cat(got, 0)[:n].cpu().numpy()

    def generate_sample(self):
        return self.generate_free_samples_array(1, free_only=False)[0]

    def generate_free_sample(self):
        return self.generate_free_samples_array(1)[0]

    def generate_free_samples(self, n):
        return [Node(id=len(self.nodes) + i, q=s) for i, s in enumerate(self.generate_free_samples_array(n))]

    def build_roadmap_n(self, n):
        raise NotImplementedError()

    def build_roadmap_time(self, t):
        raise NotImplementedError()

    def add_n_samples(self, n):
        raise NotImplementedError()

    def add_build_period(self, t):
        raise NotImplementedError()

    def update_roadmap(self):
        raise NotImplementedError()

    # ------------------------------------------------------------------ distance (:92-101)
    def distance(self, q1, q2):
        raise NotImplementedError()

    def specific_distance(self, q1, q2):
        m = self.robot_models[0]
        return m.distance(m.forward_kinematics(q1), m.forward_kinematics(q2))

    def general_distance(self, q1, q2):
        return np.linalg.norm(np.array(q1) - np.array(q2))

    # ------------------------------------------------------------------ collision (:104-169)
    def _hit(self, qs, flags):
        q = core.to_device_f64(np.atleast_2d(np.asarray(qs, dtype=np.float64)))
        return self.env.world.collide_configs(q, self.ROBOT, flags).cpu().numpy().astype(bool)

    def is_collision_free_sample(self, sample):
        return not self._hit(sample, self.FLAGS)[0]

    def robot_in_env_bounds(self, config, robot):
        robot.set_pose(config)
        return self.env.robot_in_env_bounds(robot)

    def robot_self_collision(self, config, robot):
        robot.set_pose(config)
        return self.env.robot_self_collision(robot)

    def robot_obstacle_collision(self, sample, robot, obstacle):
        robot.set_pose(sample)
        return self.env.robot_obstacle_collision(robot, obstacle)

    def is_collision_free_node(self, node):
        return self.is_collision_free_sample(node.q)

    def is_collision_free_edge(self, s, g):
        a = core.to_device_f64(np.atleast_2d(np.asarray(s, dtype=np.float64)))
        b = core.to_device_f64(np.atleast_2d(np.asarray(g, dtype=np.float64)))
        hit = self.env.world.collide_segments(a, b, self.n_steps, self.local_step, self.ROBOT, self.FLAGS)
        return not bool(hit[0].item())

    # ------------------------------------------------------------------ query: host graph search
    def find_nearest_roadmap_node(self, config):
        for node in sorted(self.nodes, key=lambda n: self.distance(config, n.q)):
            if self.is_collision_free_edge(config, node.q):
                return node
        return None

    def a_star_search(self, start_id, goal_id):
        q = self.node_id_config_dict
        g = {nid: INF for nid in self.roadmap}
        g[start_id] = 0.0
        came, done = {}, set()
        heap = [(self.distance(q[start_id], q[goal_id]), start_id)]
        while heap:
            _, cur = heapq.heappop(heap)
            if cur in done:
                continue
            done.add(cur)
            if cur == goal_id:
                path = [cur]
                while cur in came:
                    cur = came[cur]
                    path.append(cur)
                return path[::-1], g[goal_id]
            for nb in self.roadmap[cur]:
                d = g[cur] + self.distance(q[cur], q[nb])
                if d < g[nb]:
                    g[nb], came[nb] = d, cur
                    heapq.heappush(heap, (d + self.distance(q[nb], q[goal_id]), nb))
        return [], INF

    def query(self):
        start, goal = self.agents[0]['start'], self.agents[0]['goal']
        if not self.is_collision_free_sample(start) or not self.is_collision_free_sample(goal):
            return []
        s, g = self.find_nearest_roadmap_node(start), self.find_nearest_roadmap_node(goal)
        if s is None or g is None:
            return []
        path, _ = self.a_star_search(s.id, g.id)
        return [self.node_id_config_dict[i] for i in path]

    # ------------------------------------------------------------------ statistics
    @property
    def compute_nr_nodes(self):
        return len(self.nodes)

    @property
    def compute_nr_edges(self):
        return sum(len(v) for v in self.roadmap.values()) / 2

    @property
    def compute_avg_degree(self):
        return sum(len(v) for v in self.roadmap.values()) / max(len(self.roadmap), 1)

    # ------------------------------------------------------------------ shared connection passes
    def _connect_radius_causal(self, F, q64, first, r):
        """new node i (>= first) <-> every EARLIER node within r (<=) with a free edge, in the
        reference's loop order (distance_prm.py:40-52)."""
        off, idx, _ = core.radius_search(F, F.rows(first, len(self.nodes)), float(r), strict=False, exclude_self=True,
                                         causal=True, query_id0=first)
        counts = off[1:] - off[:-1]
        src = torch.repeat_interleave(torch.arange(first, len(self.nodes), device=idx.device, dtype=torch.int32), counts)
        edges = torch.stack([src, idx], 1).contiguous()
        free = self._edges_free(q64, edges)
        for a, b, ok in zip(src.cpu().numpy(), idx.cpu().numpy(), free):
            if ok:
                self.roadmap[int(a)].append(int(b))
                self.roadmap[int(b)].append(int(a))

    def _connect_knn_causal(self, F, q64, first, k):
        """new node i: earlier nodes closest first until k free edges are found
        (degree_prm.py:45-70).  Candidates come in batches that double until enough are free."""
        for i in range(max(first, 1), len(self.nodes)):
            kc = min(i, max(2 * k, 8))
            while True:
                if kc <= 58:
                    idx, _, _ = core.knn_exact(F.rows(0, i), F.rows(i, i + 1), kc, pad=min(6, i - kc),
                                               exclude_self=False)
                    cand = idx[0][idx[0] >= 0]
                else:   # obstacle-heavy neighbourhood: rank ALL earlier nodes (exact fp64 distances)
                    _, ridx, rdist = core.radius_search(F.rows(0, i), F.rows(i, i + 1), 1e300, exclude_self=False)
                    order = np.lexsort((ridx.cpu().numpy(), rdist.cpu().numpy()))
                    cand = ridx[torch.from_numpy(order).to(ridx.device)]
                    kc = i
                edges = torch.stack([torch.full_like(cand, i), cand], 1).to(torch.int32).contiguous()
                free = self._edges_free(q64, edges)
                if int(free.sum()) >= k or kc >= i:
                    break
                kc = min(i, kc * 2)
            for j, ok in zip(cand.cpu().numpy(), free):
                if len(self.roadmap[i]) >= k:
                    break
                if ok:
                    self.roadmap[i].append(int(j))
                    self.roadmap[int(j)].append(i)
