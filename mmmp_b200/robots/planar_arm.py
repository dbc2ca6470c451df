"""Planar n-link arm with the reference's interface (robots/planar_arm.py:12-61)."""
import numpy as np

from ..utils.planner_utils import Interval


class PlanarArm:
    def __init__(self, links, base_pos=np.array([0, 0])):
        self.links = np.asarray(links, dtype=float)
        self.base_pos = np.asarray(base_pos, dtype=float)
        self.n_links = self.links.shape[0]
        self.config_space = [Interval(lower=-np.pi, upper=np.pi) for _ in range(self.n_links)]
        self.q = np.zeros(self.n_links)
        self.q_d = np.zeros(self.n_links)
        self.pos = np.zeros((self.n_links, 2))

    def forward_kinematics(self, q):
        """Link end points for joint angles q (planar_arm.py:23-40)."""
        pos = np.zeros((self.n_links, 2))
        prev = self.base_pos
        for i in range(self.n_links):
            ang = q[0] if i == 0 else np.sum(q[0:i + 1])
            pos[i, :] = prev + np.array([self.links[i] * np.cos(ang), self.links[i] * np.sin(ang)])
            prev = pos[i, :]
        return pos

    def set_pose(self, q):
        self.q = q
        self.pos = self.forward_kinematics(q)

    def distance(self, pos1, pos2):
        """sum of SQUARED link end point distances (planar_arm.py:54-61)."""
        distance = 0
        for i in range(self.n_links):
            distance += np.linalg.norm(pos1[i] - pos2[i]) ** 2
        return distance
