"""CBS-PRM: conflict-based search over the per-robot roadmaps (reference
planners/prm/pybullet/decoupled/cbsprm.py:55-276).  The learning phase is DecoupledPRM's
(GPU); the search itself -- high-level constraint tree, constraint-aware A*, path
discretisation -- is host code as in the reference, with the pairwise robot-robot tests of
get_first_conflict_in_time / transition_valid batched onto the collision kernel."""
from __future__ import annotations

import time
from copy import deepcopy
from itertools import combinations

import numpy as np

from .decoupled_prm import DecoupledPRM


class Conflict:
    def __init__(self):
        self.t, self.r_id1, self.r_id2, self.q1, self.q2 = -1, -1, -1, tuple(), tuple()

    def __str__(self):
        return f"({self.t}, {self.r_id1}, {self.r_id2}, {self.q1}, {self.q2})"


class Constraint:
    def __init__(self, r_1_id, r_2_id, q2, t):
        self.r_1_id, self.r_2_id, self.q2, self.t = r_1_id, r_2_id, q2, t

    def __hash__(self):
        return hash(str(self.t) + str(self.q2))

    def __str__(self):
        return f"({self.t}, {self.q2})"


class HighLevelNode:
    def __init__(self):
        self.solution, self.discretized_solution, self.constraint_dict, self.cost = {}, {}, {}, 0

    def __eq__(self, other):
        if not isinstance(other, type(self)):
            return NotImplemented
        return self.solution == other.solution and self.cost == other.cost

    def __hash__(self):
        return hash(self.cost)

    def __lt__(self, other):
        return self.cost < other.cost


class CBSPRM(DecoupledPRM):
    def __init__(self, environment, load_roadmap, maxdist, k1=0, k2=0, build_type='n', prm_type='degree', n=0, t=0,
                 time_step=0., local_step=0., seed=0) -> None:
        super().__init__(environment, load_roadmap, maxdist, k1, k2, build_type, prm_type, n, t, time_step,
                         local_step, seed)
        self.n_ct_nodes = 0

    # ------------------------------------------------------------------ low level
    def transition_valid(self, r_id, d_to_n_1, n_1_id, n_2_id, conflict_times, constraints_from_t):
        """cbsprm.py:59-89 -- the move n1 -> n2 must not hit a constraining robot at a conflict time."""
        r_index = self.r_ids.index(r_id)
        q1 = np.array(self.node_id_q_dicts[r_index][n_1_id])
        q2 = np.array(self.node_id_q_dicts[r_index][n_2_id])
        t1 = float(d_to_n_1 / self.local_step) * self.time_step
        t2 = t1 + np.linalg.norm(q2 - q1) / self.local_step * self.time_step
        q1_t1, q2_t2 = np.append(q1, t1), np.append(q2, t2)
        it1 = np.round(np.ceil(t1 / self.time_step) * self.time_step, 1)
        it2 = np.round(np.floor(t2 / self.time_step) * self.time_step, 1)
        for t in np.round(np.arange(it1, it2 + self.time_step, self.time_step), 1):
            if t in conflict_times:
                q_at_t = self.find_q_in_q_t_interval_given_t(q1_t1, q2_t2, t)
                for constraint in constraints_from_t[t]:
                    if self.robot_robot_collision(q_at_t, constraint.q2, r_id, constraint.r_2_id):
                        return False
        return True

    def a_star_search(self, r_id, constraints):
        """cbsprm.py:217-270 (start = -1, goal = -2)."""
        r_index = self.r_ids.index(r_id)
        start_id, goal_id = -1, -2
        conflict_times = {c.t for c in constraints}
        constraints_from_t = {c.t: [] for c in constraints}
        for c in constraints:
            constraints_from_t[c.t].append(c)
        ed = self.edge_dicts[r_index]
        came_from = {}
        g_score = {nid: np.inf for nid in ed.keys()}
        f_score = dict(g_score)
        g_score[start_id] = 0
        f_score[start_id] = self.heuristic(start_id, goal_id, r_id)
        unvisited = set(ed.keys())
        while unvisited:
            current = min(unvisited, key=lambda nid: f_score[nid])
            unvisited.remove(current)
            if g_score[current] == np.inf:
                return [], np.inf
            if current == goal_id:
                return self.reconstruct_path(current, came_from), g_score[goal_id]
            for nb in ed[current]:
                if not self.transition_valid(r_id, g_score[current], current, nb, conflict_times, constraints_from_t):
                    continue
                d = self.distance(r_id, self.node_id_q_dicts[r_index][current], self.node_id_q_dicts[r_index][nb])
                if g_score[current] + d < g_score[nb]:
                    g_score[nb] = g_score[current] + d
                    f_score[nb] = g_score[nb] + self.heuristic(nb, goal_id, r_id)
                    came_from[nb] = current
        return [], np.inf

    def reconstruct_path(self, current_id, came_from):
        path = [current_id]
        while current_id in came_from:
            current_id = came_from[current_id]
            path.append(current_id)
        return path[::-1]

    def query_robot(self, r_id, constraints={}):
        path, _ = self.a_star_search(r_id, constraints)
        if not path:
            print(f"No path found for robot {r_id}.")
            return {}
        return path

    # ------------------------------------------------------------------ high level (cbsprm.py:92-215)
    def compute_solution(self, constraint_dict):
        solution = {}
        for r_id in self.r_ids:
            id_path = self.query_robot(r_id, constraint_dict[r_id])
            if not id_path:
                return {}
            solution[r_id] = id_path
        return solution

    def compute_cost(self, solution={}):
        total = 0
        for r_id, id_path in solution.items():
            d = self.node_id_q_dicts[self.r_ids.index(r_id)]
            total += sum(self.distance(r_id, d[a], d[b]) for a, b in zip(id_path[:-1], id_path[1:]))
        return total

    def get_first_conflict_in_time(self, d_solution):
        """cbsprm.py:165-183, every (t, pair) test of the sweep evaluated in one kernel launch per
        robot pair, then the first hit in (t, pair) order reported."""
        max_t = max(max(p.keys()) for p in d_solution.values())
        times = np.round(np.arange(0, max_t, self.time_step), 1)
        if len(times) == 0:
            return None
        ids = list(d_solution.keys())
        at = {}
        for r_id in ids:
            keys = np.array(list(d_solution[r_id].keys()))
            vals = list(d_solution[r_id].values())
            # min(keys, key=|x - t|): first minimum in dict order
            pick = np.abs(keys[None, :] - times[:, None]).argmin(1)
            at[r_id] = [np.asarray(vals[j], dtype=float) for j in pick]
        hits = {}
        for a, b in combinations(ids, 2):
            hits[(a, b)] = self.robot_robot_collision_batch(at[a], at[b], self.r_ids.index(a), self.r_ids.index(b))
        for ti, t in enumerate(times):
            for a, b in combinations(ids, 2):
                if hits[(a, b)][ti]:
                    c = Conflict()
                    c.t, c.r_id1, c.r_id2, c.q1, c.q2 = t, a, b, at[a][ti], at[b][ti]
                    return c
        return None

    def get_constraints_from_conflict(self, conflict):
        return {conflict.r_id1: Constraint(conflict.r_id1, conflict.r_id2, conflict.q2, conflict.t),
                conflict.r_id2: Constraint(conflict.r_id2, conflict.r_id1, conflict.q1, conflict.t)}

    def query(self):
        start = HighLevelNode()
        self.n_ct_nodes += 1
        t0 = time.perf_counter()
        for r_id in self.r_ids:
            self.add_start_goal_nodes(r_id)
            start.constraint_dict[r_id] = []
        l_t = time.perf_counter() - t0
        t0 = time.perf_counter()
        start.solution = self.compute_solution(start.constraint_dict)
        if not start.solution:
            return {}, l_t, 0
        start.discretized_solution = {r: self.discretize_path_in_time(r, start.solution[r]) for r in start.solution}
        start.cost = self.compute_cost(start.solution)
        open_set, closed_set = {start}, set()
        while open_set:
            P = min(open_set)
            open_set.remove(P)
            closed_set.add(P)
            conflict = self.get_first_conflict_in_time(P.discretized_solution)
            if conflict is None:
                print("Solution found")
                return P.solution, l_t, time.perf_counter() - t0
            constraints = self.get_constraints_from_conflict(conflict)
            for r_id in (conflict.r_id1, conflict.r_id2):
                node = deepcopy(P)
                self.n_ct_nodes += 1
                node.constraint_dict[r_id].append(constraints[r_id])
                node.solution = self.compute_solution(node.constraint_dict)
                if not node.solution:
                    continue
                node.discretized_solution = {r: self.discretize_path_in_time(r, node.solution[r])
                                             for r in node.solution}
                node.cost = self.compute_cost(node.solution)
                if node not in closed_set:
                    open_set |= {node}
        return {}, l_t, time.perf_counter() - t0
