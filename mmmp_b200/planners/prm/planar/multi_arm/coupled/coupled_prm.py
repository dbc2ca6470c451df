"""Planar composite-space PRM base (reference
planners/prm/planar/multi_arm/coupled/coupled_prm.py:19-169)."""
from __future__ import annotations

import numpy as np

from ......_lib import CHECK_BOUNDS, CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF
from ...env import LOCAL_STEP
from ...single_arm.prm import PRM


class CoupledPRM(PRM):
    ROBOT = -1
    FLAGS = CHECK_BOUNDS | CHECK_SELF | CHECK_OBSTACLES | CHECK_ROBOTS   # is_collision_free_sample :83-103

    def __init__(self, environment, seed=0) -> None:
        super().__init__(environment, local_step=LOCAL_STEP, seed=seed)
        self.config_space = self.create_composite_config_space(environment.robot_models)

    def create_composite_config_space(self, robot_models):
        space = []
        for m in robot_models:
            space.extend(m.config_space)
        return space

    def merge_configs(self, configs):
        return np.array(np.concatenate([np.array(c) for c in configs]))

    def split_config(self, config):
        out, off = [], 0
        for robot in self.robot_models:
            dim = len(robot.config_space)
            out.append(np.array(config[off:off + dim]))
            off += dim
        return out

    def specific_distance(self, q1, q2):
        a, b = self.split_config(q1), self.split_config(q2)
        return sum(r.distance(r.forward_kinematics(a[i]), r.forward_kinematics(b[i]))
                   for i, r in enumerate(self.robot_models))

    def robot_robot_collision(self, sample, robot1, robot2):
        s1, s2 = self.split_config(sample)
        robot1.set_pose(s1)
        robot2.set_pose(s2)
        return self.env.robot_robot_collision(robot1, robot2)

    def query(self):
        paths = {}
        start = np.concatenate([np.asarray(a['start'], float) for a in self.agents])
        goal = np.concatenate([np.asarray(a['goal'], float) for a in self.agents])
        if not self.is_collision_free_sample(start):
            print("Start configuration is in collision.")
            return paths
        if not self.is_collision_free_sample(goal):
            print("Goal configuration is in collision.")
            return paths
        s, g = self.find_nearest_roadmap_node(start), self.find_nearest_roadmap_node(goal)
        if s is None or g is None:
            return paths
        path, _ = self.a_star_search(s.id, g.id)
        for nid in path:
            parts = self.split_config(self.node_id_config_dict[nid])
            for i, agent in enumerate(self.agents):
                paths.setdefault(agent['name'], []).append(tuple(parts[i]))
        return paths

    def compute_combined_avg_degree(self):
        return sum(len(v) for v in self.roadmap.values()) / len(self.roadmap)

    def compute_combined_nr_nodes(self):
        return len(self.nodes)

    def compute_combined_nr_edges(self):
        return sum(len(v) for v in self.roadmap.values()) / 2
