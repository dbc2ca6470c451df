"""Planar composite-space degree PRM (reference
planners/prm/planar/multi_arm/coupled/degree_coupled_prm.py:14-98)."""
from __future__ import annotations

from ...... import core
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

    def reset_roadmap(self):
        for node in self.nodes:
            self.roadmap[node.id] = []

    def update_roadmap(self):
        return self.update_roadmap_strict_knn()

    def update_roadmap_strict_knn(self):
        q64 = self._device_nodes()
        F = self._l2(q64)
        k = min(self.n_knn, len(self.nodes) - 1)
        idx, dist, amb = core.knn_exact(F, F, k)
        free = self._edges_free(q64, core.edges_from_knn(idx)).reshape(idx.shape)
        idx = idx.cpu().numpy()
        self.last_candidates = {"idx": idx, "dist": dist.cpu().numpy(), "edge_free": free}
        self.reset_roadmap()
        for i in range(len(self.nodes)):
            for c in range(k):
                j = int(idx[i, c])
                if j >= 0 and free[i, c] and j not in self.roadmap[i]:
                    self.roadmap[i].append(j)
                    self.roadmap[j].append(i)

    def distance(self, q1, q2):
        return self.general_distance(q1, q2)
