"""Planar CBS-PRM, learning phase (reference
planners/prm/planar/multi_arm/decoupled/cbs_prm.py:69-155): per arm, every new node is linked
to every earlier node within max_edge_len (joint-space L2, <=) whose straight edge is free.
The conflict-based search of the query phase (:204-536) is host code outside the
roadmap-construction path; `robot_robot_collision` -- the predicate it is built on -- is
provided in batched form for it."""
from __future__ import annotations

import numpy as np
import torch

from ...... import core
from ......_lib import CHECK_ROBOTS
from .decoupled_prm import DecoupledPRM


class CBSPRM(DecoupledPRM):
    def __init__(self, environment, max_edge_len, timestep=0.1, build_type='n', n_nodes=100, seed=0) -> None:
        super().__init__(environment, seed=seed)
        self.max_edge_len = max_edge_len
        self.timestep = timestep
        self.starts = {i: self.agents[i]['start'] for i in self.robot_ids}
        self.goals = {i: self.agents[i]['goal'] for i in self.robot_ids}
        if build_type == 'n':
            self.build_roadmap_n(int(input("Enter the number of initial nodes per robot: ")))
        elif build_type == 't':
            raise NotImplementedError("time-boxed building is interactive in the reference")
        elif build_type == 'n_nodes':
            self.build_roadmap_n(n_nodes)
        self.collision_checks = 0
        self.n_ct_nodes = 0

    def build_roadmap_n(self, n):
        for i in self.robot_ids:
            self.add_n_samples(n, i)

    def add_n_samples(self, n, id, samples=None):
        """:100-112"""
        id = int(id)
        first = len(self.node_lists[id])
        self._register_arm(id, self.generate_free_samples_array(n, id) if samples is None else samples)
        self._connect_causal_arm(id, first)

    def _connect_causal_arm(self, id, first):
        q64 = self._device_nodes_arm(id)
        F = core.l2_features(q64)
        n = q64.shape[0]
        off, idx, _ = core.radius_search(F, F.rows(first, n), float(self.max_edge_len), strict=False, exclude_self=True,
                                         causal=True, query_id0=first)
        counts = off[1:] - off[:-1]
        src = torch.repeat_interleave(torch.arange(first, n, device=idx.device, dtype=torch.int32), counts)
        free = self._edges_free_arm(id, q64, torch.stack([src, idx], 1).contiguous())
        rm = self.roadmaps[id]
        for a, b, ok in zip(src.cpu().numpy(), idx.cpu().numpy(), free):
            if ok:
                rm[int(a)].append(int(b))
                rm[int(b)].append(int(a))

    def add_and_update(self, n):
        for i in self.robot_ids:
            self.add_n_samples(n, i)

    def update_roadmap(self):
        """:134-143 -- all ordered pairs, directed lists in ascending id."""
        for r in self.robot_ids:
            r = int(r)
            q64 = self._device_nodes_arm(r)
            F = core.l2_features(q64)
            off, idx, _ = core.radius_search(F, F, float(self.max_edge_len), strict=False, exclude_self=True)
            counts = off[1:] - off[:-1]
            src = torch.repeat_interleave(torch.arange(q64.shape[0], device=idx.device, dtype=torch.int32), counts)
            free = self._edges_free_arm(r, q64, torch.stack([src, idx], 1).contiguous())
            for node in self.node_lists[r]:
                self.roadmaps[r][node.id] = []
            for a, b, ok in zip(src.cpu().numpy(), idx.cpu().numpy(), free):
                if ok:
                    self.roadmaps[r][int(a)].append(int(b))

    def distance(self, q1, q2, robot_id):
        return self.general_distance(q1, q2)

    def robot_robot_collision(self, config1, config2, robot_id1, robot_id2):
        self.robot_models[robot_id1].set_pose(config1)
        self.robot_models[robot_id2].set_pose(config2)
        return self.env.robot_robot_collision(self.robot_models[robot_id1], self.robot_models[robot_id2])

    def find_q_in_q_t_interval_given_t(self, q1_t1, q2_t2, t):
        x = (t - q1_t1[-1]) / (q2_t2[-1] - q1_t1[-1])
        return np.round(q1_t1[:-1] + (q2_t2[:-1] - q1_t1[:-1]) * x, 2)

    def heuristic(self, node_id1, node_id2, robot_id):
        d = self.node_id_config_dict[robot_id]
        return self.distance(d[node_id1], d[node_id2], robot_id)

    def query(self):
        raise NotImplementedError("the planar CBS search is host code outside the roadmap-construction path "
                                  "(SURVEY.md 8f); use the roadmaps with the reference's own search")
