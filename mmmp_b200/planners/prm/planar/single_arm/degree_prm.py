"""Planar degree PRMs (reference planners/prm/planar/single_arm/degree_prm.py:16-211)."""
from __future__ import annotations

from ..... import core
from ..env import LOCAL_STEP
from .prm import PRM


class KNNPRM(PRM):
    """each new node is linked to its nearest EARLIER nodes (specific_distance), closest first,
    until n_knn collision-free edges exist (:45-70)."""

    def __init__(self, environment, max_edge_len, n_knn, build_type='n', n_nodes=None, seed=0):
        super().__init__(environment, local_step=LOCAL_STEP, seed=seed)
        self.max_edge_len = max_edge_len
        self.n_knn = n_knn
        if n_nodes is not None:
            self.build_roadmap_n(n_nodes)
        elif build_type == 'n':
            self.build_roadmap_n(int(input("Enter the number of initial nodes: ")))
        elif build_type == 't':
            raise NotImplementedError("time-boxed building is interactive in the reference")

    def build_roadmap_n(self, n):
        return self.add_n_samples(n)

    def add_and_update(self, n):
        self.add_n_samples(n)

    def add_n_samples(self, n, samples=None):
        first = len(self.nodes)
        self._register(self.generate_free_samples_array(n) if samples is None else samples)
        q64 = self._device_nodes()
        self._connect_knn_causal(self._fk_sq(q64), q64, first, self.n_knn)

    def distance(self, q1, q2):
        return self.specific_distance(q1, q2)


class KDTreeKNNPRM(PRM):
    """all nodes first, then KDTree.query(k+1)[1:] in joint space, symmetric insert (:171-190)."""

    def __init__(self, environment, max_edge_len, n_knn, n_nodes=None, seed=0):
        super().__init__(environment, local_step=LOCAL_STEP, seed=seed)
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
        self.reset_roadmap()
        for i in range(len(self.nodes)):
            for c in range(k):
                j = int(idx[i, c])
                if j >= 0 and free[i, c] and j not in self.roadmap[i]:
                    self.roadmap[i].append(j)
                    self.roadmap[j].append(i)

    def distance(self, q1, q2):
        return self.general_distance(q1, q2)
