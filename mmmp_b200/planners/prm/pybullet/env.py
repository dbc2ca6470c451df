"""Environment for the Panda / KUKA planners: the reference's collision predicates
(planners/prm/pybullet/env.py:15-84) answered by the CUDA sphere-model kernels instead of
pybullet.getClosestPoints.  `obstacles` are primitive descriptions (mmmp_b200.core.plane /
box / sphere / cylinder) in place of pybullet body ids.  The motion-playback half of the
reference class (env.py:86-293) is GUI code and out of scope."""
from __future__ import annotations

import numpy as np
import torch

from .... import core
from ...._lib import CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF


class Environment:
    def __init__(self, agents, obstacles):
        self.agents = agents
        self.robot_models = [agent["model"] for agent in agents]
        self.obstacles = list(obstacles)
        off, self.specs = 0, []
        for m in self.robot_models:
            self.specs.append(core.RobotSpec(m.model, m.base_matrix, off))
            off += m.model["n_joints"]
        self.world = core.SpatialWorld(self.specs, self.obstacles)
        self._obstacle_worlds = {}
        # the reference poses every robot at its GOAL, then checks (env.py:22-29)
        for i, agent in enumerate(agents):
            self.robot_models[i].set_arm_pose(agent["goal"])
        for i, agent in enumerate(agents):
            if self.robot_collision(self.robot_models[i]):
                raise ValueError(f"Agent {agent['name']} start configuration is in collision.")

    # ------------------------------------------------------------------ helpers
    def robot_index(self, robot) -> int:
        for i, m in enumerate(self.robot_models):
            if m is robot:
                return i
        raise ValueError("robot is not part of this environment")

    def _q32(self, robot) -> torch.Tensor:
        q = np.asarray(robot.get_arm_pose(), dtype=np.float64)[None, :]
        return core.pack_f32(core.to_device_f64(q))

    def reset_robots(self):
        for i, agent in enumerate(self.agents):
            self.robot_models[i].set_arm_pose(agent["start"])

    # ------------------------------------------------------------------ predicates
    def robot_collision(self, robot):
        if self.robot_self_collision(robot):
            print("Robot self collision")
            return True
        for obstacle in self.obstacles:
            if self.robot_obstacle_collision(robot, obstacle):
                print("Robot obstacle collision")
                return True
        for other_robot in self.robot_models:
            if other_robot is not robot and self.robot_robot_collision(robot, other_robot):
                print("Robot robot collision")
                return True
        return False

    def robot_self_collision(self, robot, collision_threshold=0.0):
        self._check_threshold(collision_threshold)
        r = self.robot_index(robot)
        return bool(self.world.collide_configs(self._q32(robot), r, CHECK_SELF)[0].item())

    def robot_obstacle_collision(self, robot, obstacle, collision_threshold=0.0):
        self._check_threshold(collision_threshold)
        r = self.robot_index(robot)
        key = (r, id(obstacle))
        if key not in self._obstacle_worlds:
            self._obstacle_worlds[key] = core.SpatialWorld([core.RobotSpec(robot.model, robot.base_matrix, 0)],
                                                           [obstacle])
        return bool(self._obstacle_worlds[key].collide_configs(self._q32(robot), 0, CHECK_OBSTACLES)[0].item())

    def robot_robot_collision(self, robot1, robot2, collision_threshold=0.0):
        self._check_threshold(collision_threshold)
        a, b = self.robot_index(robot1), self.robot_index(robot2)
        hit = self.world.collide_robot_pairs(self._q32(robot1), self._q32(robot2), a, b)
        return True if hit[0].item() else None      # the reference returns None, not False (env.py:75)

    @staticmethod
    def _check_threshold(t):
        if t != 0.0:
            raise ValueError("only collision_threshold = 0.0 (the value every reference caller uses) is supported")
