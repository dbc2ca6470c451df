"""Distance PRM in composite space (reference
planners/prm/pybullet/coupled/distance_coupled_prm.py:15-141)."""
from __future__ import annotations

import numpy as np
import torch

from ..... import core
from .coupled_prm import CoupledPRM


class DistanceCoupledPRM(CoupledPRM):
    """Every new node is linked to every EARLIER node within max_edge_len whose straight edge
    is collision free (add_n_samples :40-52).  The incremental loop is order dependent only
    through node ids, so it is evaluated as one causal radius search + one edge batch."""

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
        self._connect_causal(first)

    def add_and_update(self, n):
        self.add_n_samples(n)

    def _connect_causal(self, first):
        q64, q32 = self._device_nodes()
        F = core.l2_features(q64, q32)
        off, idx, _ = core.radius_search(F, F.rows(first, len(self.nodes)), float(self.max_edge_len), strict=False,
                                         exclude_self=True, causal=True, query_id0=first)
        counts = off[1:] - off[:-1]
        src = torch.repeat_interleave(torch.arange(first, len(self.nodes), device=idx.device, dtype=torch.int32), counts)
        edges = torch.stack([src, idx], 1).contiguous()
        hit = self.env.world.collide_edges(q32, edges, self.n_steps, self.local_step, -1, self.ALL)
        src, idx, free = src.cpu().numpy(), idx.cpu().numpy(), (hit == 0).cpu().numpy()
        for a, b, ok in zip(src, idx, free):      # rows ascending in new id, ids ascending inside a row
            if ok:
                self.roadmap[int(a)].append(int(b))
                self.roadmap[int(b)].append(int(a))

    def update_roadmap(self):
        """:78-85 -- rebuild all edges (i < j)"""
        for node in self.nodes:
            self.roadmap[node.id] = []
        self._connect_causal(0)

    def distance(self, q1, q2):
        return self.general_distance(q1, q2)

    def compute_total_path_length(self, paths):
        return sum(sum(self.distance(p[i], p[i + 1]) for i in range(len(p) - 1)) for p in paths.values())


class KDTreeDistanceCoupledPRM(CoupledPRM):
    """Sample everything, then link by radius (:93-141).  Reference quirk kept: the ball query is
    sorted by INDEX and its first element is dropped whether or not it is the node itself."""

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
        q64, q32 = self._device_nodes()
        F = core.l2_features(q64, q32)
        off, idx, _ = core.radius_search(F, F, float(self.max_edge_len), strict=False, exclude_self=False)
        counts = off[1:] - off[:-1]
        keep = torch.ones(idx.shape[0], dtype=torch.bool, device=idx.device)
        keep[off[:-1][counts > 0]] = False                          # indices[1:]
        src = torch.repeat_interleave(torch.arange(len(self.nodes), device=idx.device, dtype=torch.int32),
                                      (counts - (counts > 0).to(counts.dtype)))
        dst = idx[keep]
        edges = torch.stack([src, dst], 1).contiguous()
        hit = self.env.world.collide_edges(q32, edges, self.n_steps, self.local_step, -1, self.ALL)
        for node in self.nodes:
            self.roadmap[node.id] = []
        for a, b, ok in zip(src.cpu().numpy(), dst.cpu().numpy(), (hit == 0).cpu().numpy()):
            if ok:
                self.roadmap[int(a)].append(int(b))

    def distance(self, q1, q2):
        return self.general_distance(q1, q2)
