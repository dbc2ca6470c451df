"""Planar composite-space distance PRMs (reference
planners/prm/planar/multi_arm/coupled/distance_coupled_prm.py:15-141)."""
from __future__ import annotations

import torch

from ...... import core
from .coupled_prm import CoupledPRM


class DistanceCoupledPRM(CoupledPRM):
    def __init__(self, environment, max_edge_len, build_type='n', n_nodes=100, seed=0) -> None:
        super().__init__(environment, seed=seed)
        self.max_edge_len = max_edge_len
        if build_type == 'n':
            self.build_roadmap_n(int(input("Enter the number of initial nodes: ")))
        elif build_type == 't':
            raise NotImplementedError("time-boxed building is interactive in the reference")
        elif build_type == 'n_nodes':
            self.build_roadmap_n(n_nodes)

    def build_roadmap_n(self, n):
        return self.add_n_samples(n)

    def add_n_samples(self, n, samples=None):
        first = len(self.nodes)
        self._register(self.generate_free_samples_array(n) if samples is None else samples)
        q64 = self._device_nodes()
        self._connect_radius_causal(self._l2(q64), q64, first, self.max_edge_len)

    def add_and_update(self, n):
        self.add_n_samples(n)

    def update_roadmap(self):
        for node in self.nodes:
            self.roadmap[node.id] = []
        q64 = self._device_nodes()
        self._connect_radius_causal(self._l2(q64), q64, 0, self.max_edge_len)

    def distance(self, q1, q2):
        return self.general_distance(q1, q2)

    def compute_total_path_length(self, paths):
        return sum(sum(self.distance(p[i], p[i + 1]) for i in range(len(p) - 1)) for p in paths.values())


class KDTreeDistanceCoupledPRM(CoupledPRM):
    """ball query sorted by index, first element dropped (reference quirk :133-134), directed lists."""

    def __init__(self, environment, max_edge_len, n_nodes=None, seed=0):
        super().__init__(environment, seed=seed)
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
        off, idx, _ = core.radius_search(F, F, float(self.max_edge_len), strict=False, exclude_self=False)
        counts = off[1:] - off[:-1]
        keep = torch.ones(idx.shape[0], dtype=torch.bool, device=idx.device)
        keep[off[:-1][counts > 0]] = False
        src = torch.repeat_interleave(torch.arange(len(self.nodes), device=idx.device, dtype=torch.int32),
                                      counts - (counts > 0).to(counts.dtype))
        dst = idx[keep]
        free = self._edges_free(q64, torch.stack([src, dst], 1).contiguous())
        for node in self.nodes:
            self.roadmap[node.id] = []
        for a, b, ok in zip(src.cpu().numpy(), dst.cpu().numpy(), free):
            if ok:
                self.roadmap[int(a)].append(int(b))

    def distance(self, q1, q2):
        return self.general_distance(q1, q2)
