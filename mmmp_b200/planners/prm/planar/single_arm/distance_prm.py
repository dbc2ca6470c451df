"""Planar distance PRMs (reference planners/prm/planar/single_arm/distance_prm.py:16-134)."""
from __future__ import annotations

import torch

from ..... import core
from ..env import LOCAL_STEP
from .prm import PRM


class DistancePRM(PRM):
    """new node <-> every earlier node with specific_distance <= max_edge_len and a free edge."""

    def __init__(self, environment, max_edge_len, build_type='n', n_nodes=None, seed=0):
        super().__init__(environment, local_step=LOCAL_STEP, seed=seed)
        self.max_edge_len = max_edge_len
        if n_nodes is not None:
            self.build_roadmap_n(n_nodes)
        elif build_type == 'n':
            self.build_roadmap_n(int(input("Enter the number of initial nodes: ")))
        elif build_type == 't':
            raise NotImplementedError("time-boxed building is interactive in the reference")

    def build_roadmap_n(self, n):
        return self.add_n_samples(n)

    def add_n_samples(self, n, samples=None):
        first = len(self.nodes)
        self._register(self.generate_free_samples_array(n) if samples is None else samples)
        q64 = self._device_nodes()
        self._connect_radius_causal(self._fk_sq(q64), q64, first, self.max_edge_len)

    def add_and_update(self, n):
        self.add_n_samples(n)

    def update_roadmap(self):
        for node in self.nodes:
            self.roadmap[node.id] = []
        q64 = self._device_nodes()
        self._connect_radius_causal(self._fk_sq(q64), q64, 0, self.max_edge_len)

    def distance(self, q1, q2):
        return self.specific_distance(q1, q2)


class KDTreeDistancePRM(PRM):
    """all nodes first, then KDTree.query_ball_point(r) in joint space; DIRECTED lists
    (roadmap[i] only) in the tree's order (:119-131) -- here ascending id."""

    def __init__(self, environment, max_edge_len, n_nodes=None, seed=0):
        super().__init__(environment, local_step=LOCAL_STEP, seed=seed)
        self.max_edge_len = max_edge_len
        self.build_roadmap_n(int(input("Enter the number of initial nodes: ")) if n_nodes is None else n_nodes)

    def build_roadmap_n(self, n):
        self.add_and_update(n)

    def add_and_update(self, n):
        self.add_n_samples(n)
        self.update_roadmap()

    def add_n_samples(self, n, samples=None):
        self._register(self.generate_free_samples_array(n) if samples is None else samples)

    def update_roadmap(self):
        q64 = self._device_nodes()
        F = self._l2(q64)
        off, idx, _ = core.radius_search(F, F, float(self.max_edge_len), strict=False, exclude_self=True)
        counts = off[1:] - off[:-1]
        src = torch.repeat_interleave(torch.arange(len(self.nodes), device=idx.device, dtype=torch.int32), counts)
        edges = torch.stack([src, idx], 1).contiguous()
        free = self._edges_free(q64, edges)
        for node in self.nodes:
            self.roadmap[node.id] = []
        for a, b, ok in zip(src.cpu().numpy(), idx.cpu().numpy(), free):
            if ok:
                self.roadmap[int(a)].append(int(b))

    def distance(self, q1, q2):
        return self.general_distance(q1, q2)
