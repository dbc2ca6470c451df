"""Degree (k-NN) PRM in composite space (reference
planners/prm/pybullet/coupled/degree_coupled_prm.py:14-98)."""
from __future__ import annotations

import numpy as np

from ..... import core
from .coupled_prm import CoupledPRM


class KDTreeKNNCoupledPRM(CoupledPRM):
    def __init__(self, environment, max_edge_len, n_knn, n_nodes=None, seed=0):
        super().__init__(environment, seed=seed)
        self.max_edge_len = max_edge_len
        self.n_knn = n_knn
        self.build_roadmap_n(int(input("Enter the number of initial nodes: ")) if n_nodes is None else n_nodes)

    def build_roadmap_n(self, n):
        self.add_n_samples(n)
        self.update_roadmap()

    def add_and_update(self, n):
        self.add_n_samples(n)
        self.update_roadmap()

    def add_n_samples(self, n, samples=None):
        self._register(self.generate_free_samples_array(n) if samples is None else samples)

    def update_roadmap(self):
        return self.update_roadmap_strict_knn()

    def reset_roadmap(self):
        for node in self.nodes:
            self.roadmap[node.id] = []

    def update_roadmap_strict_knn(self):
        """:60-73 -- KDTree.query(q, k+1)[1:], every candidate edge checked, symmetric insert."""
        q64, q32 = self._device_nodes()
        F = core.l2_features(q64, q32)
        k = min(self.n_knn, len(self.nodes) - 1)
        idx, dist, amb = core.knn_exact(F, F, k)
        edges = core.edges_from_knn(idx)
        hit = self.env.world.collide_edges(q32, edges, self.n_steps, self.local_step, -1, self.ALL)
        free = (hit == 0).reshape(idx.shape).cpu().numpy()
        idx = idx.cpu().numpy()
        self.last_candidates = {"idx": idx, "dist": dist.cpu().numpy(), "edge_free": free}
        self.reset_roadmap()
        for i in range(len(self.nodes)):
            for c in range(k):
                j = int(idx[i, c])
                if j >= 0 and free[i, c] and j not in self.roadmap[i]:
                    self.roadmap[i].append(j)
                    self.roadmap[j].append(i)

    def update_roadmap_min_knn(self):
        """:75-90 -- ball query sorted by index, first dropped, stop at n_knn neighbours."""
        import torch
        q64, q32 = self._device_nodes()
        F = core.l2_features(q64, q32)
        off, idx, _ = core.radius_search(F, F, float(self.max_edge_len), strict=False, exclude_self=False)
        counts = off[1:] - off[:-1]
        src = torch.repeat_interleave(torch.arange(len(self.nodes), device=idx.device, dtype=torch.int32), counts)
        edges = torch.stack([src, idx], 1).contiguous()
        hit = self.env.world.collide_edges(q32, edges, self.n_steps, self.local_step, -1, self.ALL)
        off, idx, free = off.cpu().numpy(), idx.cpu().numpy(), (hit == 0).cpu().numpy()
        self.reset_roadmap()
        for i in range(len(self.nodes)):
            for e in range(off[i] + 1, off[i + 1]):
                if len(self.roadmap[i]) >= self.n_knn:
                    break
                j = int(idx[e])
                if free[e] and j not in self.roadmap[i]:
                    self.roadmap[i].append(j)
                    self.roadmap[j].append(i)

    def distance(self, q1, q2):
        return self.general_distance(q1, q2)
