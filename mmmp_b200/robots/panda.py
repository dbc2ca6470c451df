"""Franka Panda with the reference's interface (robots/panda.py:17-259), pybullet free.

Forward kinematics and the distance metrics are host float64 numpy for single
configurations (what the query phase and the reference's own callers use); the batched
versions run on the GPU through the C ABI (core.SpatialWorld.fk / fk_features)."""
from __future__ import annotations

import numpy as np

from .. import core
from .robot import Robot


class Panda(Robot):
    FRANKA_URDF = "franka_panda/panda.urdf"

    def __init__(self, fixed_base=True, base_position=(0, 0, 0), base_orientation=(0, 0, 0, 1), scale=1.0):
        super().__init__(Panda.FRANKA_URDF, fixed_base, base_position, base_orientation, scale,
                         extra_joint_limits=[(0.0, 0.04), (0.0, 0.04)])
        self.tool_link = 8
        self.arm_joints = self.get_arm_joints()
        self.gripper_joints = self.get_gripper_joints()
        self.c_space = self.get_config_space(self.joints)
        self.arm_c_space = self.get_config_space(self.arm_joints)
        self.gripper_c_space = self.get_config_space(self.gripper_joints)
        self.dimension = len(self.joints)
        self.arm_dimension = len(self.arm_joints)
        self.gripper_dimension = len(self.gripper_joints)
        self.dh_params = [dict(r, theta=0) for r in self.model["dh_modified"]]

    # getters / setters --------------------------------------------------------------
    def get_arm_joints(self):
        return self.joints[:-2]

    def get_gripper_joints(self):
        return self.joints[-2:]

    def get_arm_pose(self):
        return self.get_pose(self.arm_joints)

    def get_gripper_pose(self):
        return self.get_pose(self.gripper_joints)

    def set_arm_pose(self, pose):
        return self.set_pose(self.arm_joints, pose)

    def set_arm_in_standby(self):
        self.set_arm_pose((0, 0, 0, 0, 0, 0, 0))

    def set_arm_in_neutral(self):
        self.set_arm_pose((0, 0, 0, -1.51, 0, 1.877, 0))

    def set_gripper_pose(self, pose):
        return self.set_pose(self.gripper_joints, pose)

    def set_gripper_open(self):
        self.set_gripper_pose((0.04, 0.04))

    def set_gripper_close(self):
        self.set_gripper_pose((0, 0))

    def execute_arm_motion(self, arm_path):
        for q in arm_path:
            self.set_arm_pose(q)

    def execute_gripper_motion(self, gripper_path):
        for q in gripper_path:
            self.set_gripper_pose(q)

    # forward kinematics (panda.py:111-178) --------------------------------------------
    def modified_homogeneous_transform(self, a, d, alpha, theta):
        ct, st, ca, sa = np.cos(theta), np.sin(theta), np.cos(alpha), np.sin(alpha)
        return np.array([[ct, -st, 0, a], [st * ca, ct * ca, -sa, -sa * d], [st * sa, ct * sa, ca, ca * d],
                         [0, 0, 0, 1]])

    def _frames(self, q):
        T = core.base_matrix(self.base_position, self.base_orientation)
        out = []
        for i, p in enumerate(self.dh_params):
            theta = q[i] if i < 7 else p["theta"]
            T = np.matmul(T, self.modified_homogeneous_transform(p["a"], p["d"], p["alpha"], theta))
            out.append(T)
        return out

    def forward_kinematics(self, q):
        return self._frames(q)[-1]

    def position_from_fk(self, q):
        return self.forward_kinematics(q)[:3, 3]

    def positions_from_fk(self, q):
        return [T[:3, 3] for T in self._frames(q)]

    def orientation_from_fk(self, q):
        return self.forward_kinematics(q)[:3, :3]

    # inverse kinematics (panda.py:204-225).  The reference hands these to
    # pybullet.calculateInverseKinematics (utils/ik_utils.py:21-27); here they run the batched
    # damped-least-squares kernel (mmmp_ik_dls) on a private single-robot world.  Parity with
    # pybullet's solver is unpinned.
    def _ik_world(self):
        if getattr(self, "_ik_w", None) is None:
            base = core.base_matrix(self.base_position, self.base_orientation)
            self._ik_w = core.SpatialWorld([core.RobotSpec(self.model, base, 0)], [])
        return self._ik_w

    def solve_closest_arm_ik(self, targetPose, initialJointPositions):
        import torch
        pos, quat = targetPose
        dev = core._require_cuda()
        q, conv, _ = self._ik_world().ik(torch.tensor([list(pos)], dtype=torch.float64, device=dev),
                                         torch.tensor([list(quat)], dtype=torch.float64, device=dev),
                                         torch.tensor([list(initialJointPositions)[:7]], dtype=torch.float64, device=dev))
        return tuple(q[0].cpu().tolist())

    def solve_arm_ik(self, targetPose):
        return self.solve_closest_arm_ik(targetPose, self.get_arm_pose())

    def solve_ik(self, target_position, target_orientation, default_joint_positions):
        return self.solve_closest_arm_ik((target_position, target_orientation), default_joint_positions)

    # distance metrics (panda.py:227-259) -----------------------------------------------
    def distance_metric(self, q1, q2):
        return self.task_space_pose_distance(q1, q2)

    def angular_distance_metric(self, q1, q2):
        return np.linalg.norm(np.array(q1) - np.array(q2))

    def euclidean_distance_metric(self, q1, q2):
        return np.linalg.norm(np.array(self.positions_from_fk(q1)) - np.array(self.positions_from_fk(q2)))

    def task_space_pose_distance(self, q1, q2):
        distance = 0
        for a, b in zip(self.positions_from_fk(q1), self.positions_from_fk(q2)):
            distance += np.linalg.norm(a - b)
        return distance

    def task_space_position_distance(self, q1, q2):
        return np.linalg.norm(np.array(self.position_from_fk(q1)) - np.array(self.position_from_fk(q2)))

    def task_space_orientation_distance(self, q1, q2):
        R = np.dot(np.array(self.orientation_from_fk(q1)).T, np.array(self.orientation_from_fk(q2)))
        return np.arccos((min(3, max(-1, np.trace(R))) - 1) / 2)
