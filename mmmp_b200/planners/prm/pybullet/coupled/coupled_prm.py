"""Composite-space PRM base (reference planners/prm/pybullet/coupled/coupled_prm.py:22-187):
one roadmap in the concatenated joint space of all robots; a composite configuration is
free when every arm is free of self collision, of every other arm and of the obstacles.
Collision predicates and neighbour search run on the GPU; graph search (A*/Dijkstra over
`roadmap`) stays on the host."""
from __future__ import annotations

import heapq

import numpy as np
import torch

from ..... import core
from ....._lib import CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF
from .....robots.panda import Panda
from ..utils import INF, Node


class CoupledPRM:
    def __init__(self, environment, time_step=0.1, local_step=0.05, seed=0) -> None:
        self.env = environment
        self.agents = environment.agents
        self.robot_models = environment.robot_models
        self.obstacles = environment.obstacles
        self.c_space = self.create_composite_c_space(environment.robot_models)
        self.time_step = time_step
        self.local_step = local_step
        self.nodes = []
        self.roadmap = {}
        self.node_q_id_dict = {}
        self.node_id_q_dict = {}
        self.seed = int(seed)
        self._draws = 0

    # ------------------------------------------------------------------ composite space (:45-75)
    def create_composite_c_space(self, robot_models):
        space = []
        for m in robot_models:
            space.extend(m.arm_c_space if isinstance(m, Panda) else m.c_space)
        return space

    def merge_configs(self, configs):
        return np.array(np.concatenate([np.array(c) for c in configs]))

    def split_config(self, config):
        out, off = [], 0
        for robot in self.robot_models:
            dim = len(robot.arm_c_space) if isinstance(robot, Panda) else len(robot.c_space)
            out.append(np.array(config[off:off + dim]))
            off += dim
        return out

    def update_node_id_q_dict(self, nodes):
        for node in nodes:
            self.node_id_q_dict[node.id] = node.q

    def update_node_q_id_dict(self, nodes):
        for node in nodes:
            self.node_q_id_dict[node.q] = node.id

    # ------------------------------------------------------------------ device helpers
    @property
    def n_steps(self) -> int:
        return len(np.arange(0, 1, self.local_step))

    ALL = CHECK_SELF | CHECK_OBSTACLES | CHECK_ROBOTS

    def _device_nodes(self):
        q64 = core.to_device_f64(np.array([n.q for n in self.nodes], dtype=np.float64).reshape(-1, len(self.c_space)))
        return q64, core.pack_f32(q64)

    def _register(self, samples):
        new = []
        for s in samples:
            node = Node(id=len(self.nodes), q=s)
            self.nodes.append(node)
            self.roadmap[node.id] = []
            self.node_q_id_dict[node.q] = node.id
            self.node_id_q_dict[node.id] = node.q
            new.append(node)
        return new

    # ------------------------------------------------------------------ sampling (:88-109)
    def generate_sample(self):
        return self.generate_free_samples_array(1, free_only=False)[0]

    def generate_free_sample(self):
        return self.generate_free_samples_array(1)[0]

    def generate_free_samples_array(self, n, free_only=True):
        lo = np.array([iv.lower for iv in self.c_space])
        hi = np.array([iv.upper for iv in self.c_space])
        got, have, rate = [], 0, 0.3
        while have < n:
            m = n if not free_only else int(np.ceil((n - have) / rate * 1.1)) + 32
            q64, q32 = core.sample_uniform(lo, hi, m, self.seed, self._draws)
            self._draws += m
            if free_only:
                hit = self.env.world.collide_configs(q32, -1, self.ALL)
                q64 = core.gather_rows(q64, core.compact_free(hit))
                rate = max(q64.shape[0] / m, 1e-3)
            got.append(q64)
            have += q64.shape[0]
        return torch.cat(got, 0)[:n].cpu().numpy()

    def generate_free_samples(self, n):
        q = self.generate_free_samples_array(n)
        return [Node(id=len(self.nodes) + i, q=s) for i, s in enumerate(q)]

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

    # ------------------------------------------------------------------ distance (:127-140)
    def distance(self, q1, q2):
        raise NotImplementedError()

    def general_distance(self, q1, q2):
        return np.linalg.norm(np.array(q1) - np.array(q2))

    def specific_distance(self, q1, q2):
        a, b = self.split_config(q1), self.split_config(q2)
        return sum(robot.distance_metric(a[i], b[i]) for i, robot in enumerate(self.robot_models))

    # ------------------------------------------------------------------ local planner (:143-187)
    def _hit(self, qs, flags, robot=-1):
        q32 = core.pack_f32(core.to_device_f64(np.atleast_2d(np.asarray(qs, dtype=np.float64))))
        return self.env.world.collide_configs(q32, robot, flags).cpu().numpy().astype(bool)

    def is_collision_free_sample(self, sample):
        return not self._hit(sample, self.ALL)[0]

    def robot_self_collision(self, q, robot):
        robot.set_arm_pose(q)
        return bool(self._hit(q, CHECK_SELF, self.env.robot_index(robot))[0])

    def robot_robot_collision(self, sample, robot1, robot2):
        s1, s2 = self.split_config(sample) if len(self.robot_models) == 2 else (sample[:len(sample) // 2],
                                                                                 sample[len(sample) // 2:])
        robot1.set_arm_pose(s1)
        robot2.set_arm_pose(s2)
        return self.env.robot_robot_collision(robot1, robot2)

    def robot_obstacle_collision(self, q, robot, obstacle):
        robot.set_arm_pose(q)
        return self.env.robot_obstacle_collision(robot, obstacle)

    def is_collision_free_node(self, node):
        return self.is_collision_free_sample(node.q)

    def edges_free(self, qa, qb):
        a = core.pack_f32(core.to_device_f64(np.atleast_2d(np.asarray(qa, dtype=np.float64))))
        b = core.pack_f32(core.to_device_f64(np.atleast_2d(np.asarray(qb, dtype=np.float64))))
        hit = self.env.world.collide_segments(a, b, self.n_steps, self.local_step, -1, self.ALL)
        return (hit == 0).cpu().numpy()

    def is_collision_free_edge(self, s, g):
        return bool(self.edges_free(s, g)[0])

    # ------------------------------------------------------------------ query (:190-330): host graph search
    def find_nearest_roadmap_node(self, config):
        best, best_d = None, float('inf')
        order = sorted(self.nodes, key=lambda n: self.distance(config, n.q))
        for node in order:
            d = self.distance(config, node.q)
            if d < best_d and self.is_collision_free_edge(config, node.q):
                return node
        return best

    def a_star_search(self, start_id, goal_id):
        g = {nid: INF for nid in self.roadmap}
        g[start_id] = 0.0
        came = {}
        heap = [(self.distance(self.node_id_q_dict[start_id], self.node_id_q_dict[goal_id]), start_id)]
        done = set()
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
                d = g[cur] + self.distance(self.node_id_q_dict[cur], self.node_id_q_dict[nb])
                if d < g[nb]:
                    g[nb] = d
                    came[nb] = cur
                    heapq.heappush(heap, (d + self.distance(self.node_id_q_dict[nb], self.node_id_q_dict[goal_id]), nb))
        return [], INF

    def query(self):
        paths = {}
        start = np.concatenate([np.asarray(a['start'], float) for a in self.agents])
        goal = np.concatenate([np.asarray(a['goal'], float) for a in self.agents])
        if not self.is_collision_free_sample(start):
            print("Start configuration is in collision.")
            return paths
        if not self.is_collision_free_sample(goal):
            print("Goal configuration is in collision.")
            return paths
        s, g = self.find_nearest_roadmap_node(start), self.find_nearest_roadmap_node(goal)
        if s is None or g is None:
            return paths
        path, _ = self.a_star_search(s.id, g.id)
        for nid in path:
            parts = self.split_config(self.node_id_q_dict[nid])
            for i, agent in enumerate(self.agents):
                paths.setdefault(agent['name'], []).append(tuple(parts[i]))
        return paths

    # ------------------------------------------------------------------ statistics
    def compute_combined_avg_degree(self):
        return sum(len(nb) for nb in self.roadmap.values()) / len(self.roadmap)

    def compute_combined_nr_nodes(self):
        return len(self.nodes)

    def compute_combined_nr_edges(self):
        return sum(len(nb) for nb in self.roadmap.values()) / 2
