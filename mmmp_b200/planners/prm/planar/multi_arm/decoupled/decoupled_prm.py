"""Planar decoupled PRM base (reference
planners/prm/planar/multi_arm/decoupled/decoupled_prm.py:20-152): one roadmap per arm; a
configuration of arm r is free when the arm is inside the map and clear of the obstacles
(self collision is commented out in the reference, :131-133)."""
from __future__ import annotations

import numpy as np
import torch

from ...... import core
from ......_lib import CHECK_BOUNDS, CHECK_OBSTACLES
from ...env import LOCAL_STEP
from ...single_arm.prm import PRM
from ...utils import Node


class DecoupledPRM(PRM):
    FLAGS = CHECK_BOUNDS | CHECK_OBSTACLES

    def __init__(self, environment, seed=0) -> None:
        super().__init__(environment, local_step=LOCAL_STEP, seed=seed)
        n = len(environment.robot_models)
        self.robot_ids = np.linspace(0, n - 1, n, dtype=int)
        self.config_spaces = [m.config_space for m in environment.robot_models]
        self.roadmaps = [{} for _ in range(n)]
        self.node_lists = [[] for _ in range(n)]
        self.node_id_config_dict = [{} for _ in range(n)]
        self.node_config_id_dict = [{} for _ in range(n)]
        self._draws = [0] * n

    # ------------------------------------------------------------------ sampling (:77-98)
    def generate_free_samples_array(self, n, robot_id, free_only=True):
        cs = self.config_spaces[robot_id]
        lo, hi = np.array([iv.lower for iv in cs]), np.array([iv.upper for iv in cs])
        got, have, rate = [], 0, 0.5
        while have < n:
            m = n if not free_only else int(np.ceil((n - have) / rate * 1.1)) + 32
            q64, _ = core.sample_uniform(lo, hi, m, self.seed + 7919 * int(robot_id), self._draws[robot_id])
            self._draws[robot_id] += m
            if free_only:
                hit = self.env.world.collide_configs(q64, int(robot_id), self.FLAGS)
                q64 = core.gather_rows(q64, core.compact_free(hit))
                rate = max(q64.shape[0] / m, 1e-3)
            got.append(q64)
            have += q64.shape[0]
        return torch.cat(got, 0)[:n].cpu().numpy()

    def generate_sample(self, config_space):
        lo, hi = [iv.lower for iv in config_space], [iv.upper for iv in config_space]
        q64, _ = core.sample_uniform(lo, hi, 1, self.seed, sum(self._draws) + 1)
        return q64.cpu().numpy()[0]

    def generate_free_sample(self, robot_id):
        return self.generate_free_samples_array(1, robot_id)[0]

    def generate_free_samples(self, n, robot_id):
        base = len(self.node_lists[robot_id])
        return [Node(id=base + i, q=s) for i, s in enumerate(self.generate_free_samples_array(n, robot_id))]

    # ------------------------------------------------------------------ per-arm helpers
    def _register_arm(self, robot_id, samples):
        for s in samples:
            node = Node(id=len(self.node_lists[robot_id]), q=s)
            self.node_lists[robot_id].append(node)
            self.roadmaps[robot_id][node.id] = []
            self.node_config_id_dict[robot_id][node.q] = node.id
            self.node_id_config_dict[robot_id][node.id] = node.q

    def _device_nodes_arm(self, robot_id):
        q = np.array([n.q for n in self.node_lists[robot_id]], dtype=np.float64)
        return core.to_device_f64(q.reshape(-1, len(self.config_spaces[robot_id])))

    def _edges_free_arm(self, robot_id, q64, edges):
        hit = self.env.world.collide_edges(q64, edges, self.n_steps, self.local_step, int(robot_id), self.FLAGS)
        return (hit == 0).cpu().numpy()

    def specific_distance(self, q1, q2, robot_id):
        m = self.robot_models[robot_id]
        return m.distance(m.forward_kinematics(q1), m.forward_kinematics(q2))

    # ------------------------------------------------------------------ collision (:126-152)
    def is_collision_free_sample(self, sample, robot_id):
        q = core.to_device_f64(np.atleast_2d(np.asarray(sample, dtype=np.float64)))
        return not bool(self.env.world.collide_configs(q, int(robot_id), self.FLAGS)[0].item())

    def is_collision_free_node(self, node, robot_id):
        return self.is_collision_free_sample(node.q, robot_id)

    def is_collision_free_edge(self, q1, q2, robot_id):
        a = core.to_device_f64(np.atleast_2d(np.asarray(q1, dtype=np.float64)))
        b = core.to_device_f64(np.atleast_2d(np.asarray(q2, dtype=np.float64)))
        hit = self.env.world.collide_segments(a, b, self.n_steps, self.local_step, int(robot_id), self.FLAGS)
        return not bool(hit[0].item())

    def find_nearest_roadmap_node(self, config, robot_id):
        for node in sorted(self.node_lists[robot_id], key=lambda n: self.distance(config, n.q, robot_id)):
            if self.is_collision_free_edge(config, node.q, robot_id):
                return node
        return None

    # ------------------------------------------------------------------ statistics
    def compute_combined_avg_degree(self):
        return (sum(len(v) for rm in self.roadmaps for v in rm.values()) /
                max(sum(len(rm) for rm in self.roadmaps), 1))

    def compute_combined_nr_nodes(self):
        return sum(len(nl) for nl in self.node_lists)

    def compute_nr_edges(self, robot_id):
        return sum(len(v) for v in self.roadmaps[robot_id].values()) / 2

    def compute_combined_nr_edges(self):
        return sum(self.compute_nr_edges(r) for r in range(len(self.roadmaps)))
