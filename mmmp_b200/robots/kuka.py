"""KUKA iiwa14 (sphere collision model shipped by the reference:
models/kuka_iiwa/urdf/iiwa14_spheres_collision.urdf).  The reference only loads it
(tests/models/test_kuka_iiwa.py:36-46); FK is the generic URDF chain."""
import numpy as np

from .. import core
from .robot import Robot


class Kuka(Robot):
    KUKA_URDF = "kuka_iiwa/urdf/iiwa14_spheres_collision.urdf"

    def __init__(self, fixed_base=True, base_position=(0, 0, 0), base_orientation=(0, 0, 0, 1), scale=1.0):
        super().__init__(Kuka.KUKA_URDF, fixed_base, base_position, base_orientation, scale)
        self.arm_joints = self.joints
        self.arm_c_space = self.c_space = self.config_space
        self.arm_dimension = len(self.arm_joints)

    def set_arm_pose(self, pose):
        return self.set_pose(self.arm_joints, pose)

    def get_arm_pose(self):
        return self.get_pose(self.arm_joints)

    def positions_from_fk(self, q):
        T = self.base_matrix
        out = []
        org = core.model_origins(self.model)
        for j, O in enumerate(org):
            T = T @ O
            if j < len(q):
                c, s = np.cos(q[j]), np.sin(q[j])
                T = T @ np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
            out.append(T[:3, 3])
        return out

    def distance_metric(self, q1, q2):
        return sum(np.linalg.norm(a - b) for a, b in zip(self.positions_from_fk(q1), self.positions_from_fk(q2)))
