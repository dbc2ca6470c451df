"""Planar environment (reference planners/prm/planar/env.py:21-229): map bounds, planar arms,
Rectangle / Circle obstacles.  Single-configuration predicates keep the reference's
signatures; all of them are answered by the fp64 planar kernels (mmmp_b200/csrc/planar.cu),
and the planners use the batched forms through `self.world`."""
from __future__ import annotations

import numpy as np

from .... import core
from ...._lib import CHECK_BOUNDS, CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF
from ....utils.shapes import Circle, Rectangle, is_circle, is_rectangle  # noqa: F401

LOCAL_STEP = 0.05  # [rad] step of the local planner (env.py:19)


def obstacle_desc(ob) -> dict:
    if is_rectangle(ob):
        x, y, w, h = ob.get_bbox().bounds      # env.py:170: vertices rebuilt as x + w, y + h
        return {"type": "rect", "x0": x, "y0": y, "x1": x + w, "y1": y + h}
    if is_circle(ob):
        return {"type": "circle", "cx": float(ob.center[0]), "cy": float(ob.center[1]), "r": float(ob.radius)}
    raise ValueError("obstacle must be a Rectangle or a Circle")


class Environment:
    def __init__(self, map_dimensions, agents, obstacles):
        self.map_dimensions = map_dimensions
        self.map_center = [0, 0]
        self.map_lower_bounds = [-map_dimensions[0] / 2, -map_dimensions[1] / 2]
        self.map_upper_bounds = [map_dimensions[0] / 2, map_dimensions[1] / 2]
        self.map_lower_x_bound, self.map_lower_y_bound = self.map_lower_bounds
        self.map_upper_x_bound, self.map_upper_y_bound = self.map_upper_bounds
        self.agents = agents
        self.robot_models = [agent["model"] for agent in agents]
        self.obstacles = obstacles
        for obstacle in obstacles:
            if not self.obstacle_in_env_bounds(obstacle):
                raise ValueError("Obstacle is out of environment bounds.")
        arms = [{"links": list(m.links), "base": tuple(m.base_pos)} for m in self.robot_models]
        bounds = (self.map_lower_x_bound, self.map_upper_x_bound, self.map_lower_y_bound, self.map_upper_y_bound)
        self.world = core.PlanarWorld(arms, bounds, [obstacle_desc(o) for o in obstacles])
        self._obstacle_worlds = {}
        self._posed = []
        for i, agent in enumerate(agents):
            model = self.robot_models[i]
            model.set_pose(agent["start"])
            # (the reference tests against the not-yet-posed arms' all-zero link arrays here;
            #  only arms that already have a pose are tested)
            if self.robot_collision(model, others=list(self._posed)):
                raise ValueError(f"Agent {agent['name']} start configuration is in collision.")
            self._posed.append(model)

    # ------------------------------------------------------------------ bounds (env.py:50-58,220-229)
    def point_in_env_bounds(self, point):
        return (self.map_lower_x_bound <= point[0] <= self.map_upper_x_bound and
                self.map_lower_y_bound <= point[1] <= self.map_upper_y_bound)

    def obstacle_in_env_bounds(self, obstacle):
        if is_rectangle(obstacle):
            p = obstacle.get_bbox().get_points()
            return self.point_in_env_bounds(p[0]) and self.point_in_env_bounds(p[1])
        if is_circle(obstacle):
            c, r = obstacle.center, obstacle.radius
            return all(self.point_in_env_bounds(c + np.array(d)) for d in ([-r, 0], [0, -r], [r, 0], [0, r]))

    def robot_index(self, robot) -> int:
        for i, m in enumerate(self.robot_models):
            if m is robot:
                return i
        raise ValueError("robot is not part of this environment")

    def _composite(self):
        return core.to_device_f64(np.concatenate([np.asarray(m.q, dtype=float) for m in self.robot_models])[None, :])

    def _single(self, robot, flags):
        q = core.to_device_f64(np.asarray(robot.q, dtype=float)[None, :])
        return bool(self.world.collide_configs(q, self.robot_index(robot), flags)[0].item())

    # ------------------------------------------------------------------ predicates
    def robot_in_env_bounds(self, robot):
        return not self._single(robot, CHECK_BOUNDS)

    def robot_self_collision(self, robot):
        return self._single(robot, CHECK_SELF)

    def robot_obstacle_collision(self, robot, obstacle):
        r = self.robot_index(robot)
        key = (r, id(obstacle))
        if key not in self._obstacle_worlds:
            m = self.robot_models[r]
            self._obstacle_worlds[key] = core.PlanarWorld([{"links": list(m.links), "base": tuple(m.base_pos)}],
                                                          (-1e300, 1e300, -1e300, 1e300), [obstacle_desc(obstacle)])
        q = core.to_device_f64(np.asarray(robot.q, dtype=float)[None, :])
        return bool(self._obstacle_worlds[key].collide_configs(q, 0, CHECK_OBSTACLES)[0].item())

    def robot_robot_collision(self, robot1, robot2):
        a, b = self.robot_index(robot1), self.robot_index(robot2)
        key = ("rr", a, b)
        if key not in self._obstacle_worlds:
            arms = [{"links": list(m.links), "base": tuple(m.base_pos)} for m in (robot1, robot2)]
            self._obstacle_worlds[key] = core.PlanarWorld(arms, (-1e300, 1e300, -1e300, 1e300), [])
        q = core.to_device_f64(np.concatenate([np.asarray(robot1.q, float), np.asarray(robot2.q, float)])[None, :])
        return bool(self._obstacle_worlds[key].collide_configs(q, -1, CHECK_ROBOTS)[0].item())

    def robot_collision(self, robot, others=None):
        """env.py:61-76 (self collision is commented out in the reference)."""
        others = self.robot_models if others is None else others
        if not self.robot_in_env_bounds(robot):
            print("Robot out of environment bounds")
            return True, "Robot out of environment bounds"
        for obstacle in self.obstacles:
            if self.robot_obstacle_collision(robot, obstacle):
                print("Robot obstacle collision")
                return True, "Robot obstacle collision"
        for other in others:
            if other is not robot and self.robot_robot_collision(robot, other):
                print("Robot robot collision")
                return True, "Robot robot collision"
        return False
