"""Prioritized PRM (reference planners/prm/pybullet/decoupled/prioritized_prm.py:13-167):
robots plan one after the other in priority order on their own roadmaps, each avoiding the
space-time paths of the robots planned before it.  Learning phase = DecoupledPRM (GPU);
the per-edge robot-robot sweep of robot_robot_collision_on_edge is one batched kernel call."""
from __future__ import annotations

import time
from collections import deque

import numpy as np

from ..utils import INF
from .decoupled_prm import DecoupledPRM


class PrioritizedPRM(DecoupledPRM):
    def __init__(self, environment, load_roadmap, maxdist=0, k1=0, k2=0, build_type='kdtree', prm_type='distance',
                 n=0, t=0., time_step=0., local_step=0., seed=0) -> None:
        super().__init__(environment, load_roadmap, maxdist, k1, k2, build_type, prm_type, n, t, time_step,
                         local_step, seed)

    def robot_robot_collision_on_edge(self, current_n_id, neighbor_n_id, r_id, higher_priority_robot_paths, start_time):
        """:17-49"""
        r_index = self.r_ids.index(r_id)
        a = np.array(self.node_id_q_dicts[r_index][current_n_id])
        b = np.array(self.node_id_q_dicts[r_index][neighbor_n_id])
        dt = np.linalg.norm(a - b) / self.local_step * self.time_step
        a_t, b_t = np.append(a, start_time), np.append(b, start_time + dt)
        t0 = np.round(np.ceil(start_time / self.time_step) * self.time_step, 1)
        t1 = np.round(np.floor((start_time + dt) / self.time_step) * self.time_step, 1)
        times = [np.round(t, 1) for t in np.arange(t0, t1 + self.time_step, self.time_step)]
        if not times or not higher_priority_robot_paths:
            return False
        mine = [self.find_q_in_q_t_interval_given_t(a_t, b_t, t) for t in times]
        for other_id, other_path in higher_priority_robot_paths.items():
            keys = np.array(list(other_path.keys()))
            vals = list(other_path.values())
            pick = np.abs(keys[None, :] - np.array(times)[:, None]).argmin(1)
            theirs = [np.asarray(vals[j], dtype=float) for j in pick]
            if self.robot_robot_collision_batch(mine, theirs, r_index, self.r_ids.index(other_id)).any():
                return True
        return False

    def query(self, priorities=None):
        if priorities is None:
            priorities = [i for i in range(1, len(self.r_ids) + 1)]
        assert len(priorities) == len(self.r_ids), "Number of priorities should match the number of robots."
        paths, st_paths, total_l_t, total_q_t = {}, {}, 0, 0
        for priority_index in np.argsort(priorities)[::-1]:
            r_id = self.r_ids[priority_index]
            res = self.query_robot(r_id, higher_priority_robot_paths=st_paths)
            if not res or not res[1]:
                return {}
            id_path, st_path, l_t, q_t = res
            paths[r_id] = id_path
            st_paths[r_id] = st_path
            total_l_t += l_t
            total_q_t += q_t
        return paths, total_l_t, total_q_t

    def query_robot(self, r_id, higher_priority_robot_paths):
        r_index = self.r_ids.index(r_id)
        if not self.is_collision_free_sample(self.agents[r_index]['start'], r_id):
            print(f"Start configuration for robot {r_id} is in collision.")
            return []
        if not self.is_collision_free_sample(self.agents[r_index]['goal'], r_id):
            print(f"Goal configuration for robot {r_id} is in collision.")
            return []
        t0 = time.perf_counter()
        self.add_start_goal_nodes(r_id)
        l_t = time.perf_counter() - t0
        t0 = time.perf_counter()
        path, _ = self.a_star_search(r_id, higher_priority_robot_paths)
        q_t = time.perf_counter() - t0
        if not path:
            print(f"No path found for robot {r_id}.")
            return {}
        return path, self.discretize_path_in_time(r_id, path), l_t, q_t

    def a_star_search(self, r_id, higher_priority_robot_paths):
        """:113-167 (start = -1, goal = -2)"""
        r_index = self.r_ids.index(r_id)
        start_id, goal_id = -1, -2
        ed = self.edge_dicts[r_index]
        dist = {nid: INF for nid in ed.keys()}
        est = dict(dist)
        prev = {nid: None for nid in ed.keys()}
        dist[start_id] = 0
        est[start_id] = self.heuristic(start_id, goal_id, r_id)
        unvisited = set(ed.keys())
        while unvisited:
            cur = min(unvisited, key=lambda nid: est[nid])
            unvisited.remove(cur)
            if dist[cur] == INF or cur == goal_id:
                break
            t = float(dist[cur] / self.local_step) * self.time_step
            for nb in ed[cur]:
                if self.robot_robot_collision_on_edge(cur, nb, r_id, higher_priority_robot_paths, t):
                    continue
                d = self.distance(r_id, self.node_id_q_dicts[r_index][nb], self.node_id_q_dicts[r_index][cur])
                if dist[cur] + d < dist[nb]:
                    dist[nb] = dist[cur] + d
                    est[nb] = dist[nb] + self.heuristic(nb, goal_id, r_id)
                    prev[nb] = cur
        path = deque()
        cur = goal_id
        if prev[cur] is None:
            return list(path), INF
        while prev[cur] is not None:
            path.appendleft(cur)
            cur = prev[cur]
        path.appendleft(start_id)
        return list(path), dist[goal_id]
