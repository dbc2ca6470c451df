"""Robot base class with the reference's interface (robots/robot.py:7-36) and no pybullet:
the URDF is resolved to a fitted collision / kinematic model (mmmp_b200/geometry/*.json)
and the "simulation state" is just the current joint vector held in Python."""
from __future__ import annotations

import itertools

import numpy as np

from .. import core
from ..utils.planner_utils import intervals_from_limits

_ids = itertools.count()

# URDF path (suffix) -> geometry model
MODELS = {
    "franka_panda/panda.urdf": "panda_spheres",
    "panda_arm_hand.urdf": "panda_spheres",
    "iiwa14_spheres_collision.urdf": "kuka_iiwa14_spheres",
    "iiwa14_primitive_collision.urdf": "kuka_iiwa14_spheres",
    "iiwa14_polytope_collision.urdf": "kuka_iiwa14_spheres",
}


def resolve_model(urdf_path: str) -> dict:
    if not urdf_path.endswith(".urdf"):
        raise ValueError(urdf_path)          # robots/robot.py:25
    for suffix, name in MODELS.items():
        if urdf_path.endswith(suffix):
            return core.load_model(name)
    raise ValueError(f"no fitted collision model for {urdf_path}")


class Robot:
    def __init__(self, urdf_path, fixed_base=False, base_position=(0, 0, 0), base_orientation=(0, 0, 0, 1),
                 scale=1.0, extra_joint_limits=()):
        self.urdf_path = urdf_path
        self.fixed_base = fixed_base
        self.base_position = base_position
        self.base_orientation = base_orientation
        self.scale = scale
        if scale != 1.0:
            raise ValueError("globalScaling != 1 is not supported by the fitted models")
        self.model = resolve_model(urdf_path)
        self.r_id = self.load()
        limits = list(self.model["joint_limits"]) + list(extra_joint_limits)
        self._limits = limits
        self.joints = list(range(len(limits)))
        self._q = np.zeros(len(limits))
        self.config_space = self.get_config_space(self.joints)
        self.dimension = len(self.joints)

    def load(self):
        return next(_ids)

    @property
    def base_matrix(self) -> np.ndarray:
        return core.base_matrix(self.base_position, self.base_orientation)

    def get_config_space(self, joints):
        return intervals_from_limits([self._limits[j] for j in joints])

    def set_pose(self, joints, pose):
        for j, v in zip(joints, pose):
            self._q[j] = float(v)

    def get_pose(self, joints):
        return tuple(float(self._q[j]) for j in joints)

    def execute_motion(self, path):
        for q in path:
            self.set_pose(self.joints, q)
