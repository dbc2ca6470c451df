"""Roadmap data model of the reference (planners/prm/pybullet/utils.py:13-58): nodes carry
q rounded to 2 decimals and order by id (the heap tie-break of the neighbour loops)."""
from ....utils.planner_utils import Interval

INF = float("inf")


class Node:
    __slots__ = ("id", "q")

    def __init__(self, id, q):
        self.id = id
        self.q = tuple([round(c, 2) for c in q])

    def __eq__(self, other):
        return self.id == other.id

    def __lt__(self, other):
        return self.id < other.id

    def __hash__(self):
        return hash(str(self.id) + str(self.q))

    def __str__(self):
        return str((self.id, self.q))


class Edge:
    def __init__(self, node1, node2):
        self.node1, self.node2 = node1, node2

    def __eq__(self, other):
        return self.node1 == other.node1 and self.node2 == other.node2

    def __hash__(self):
        return hash(str(self.node1) + str(self.node2))

    def __str__(self):
        return str((self.node1, self.node2))


class State:
    def __init__(self, t, node):
        self.t = t
        self.q = node.q

    def __eq__(self, other):
        return self.t == other.t and self.q == other.q

    def __hash__(self):
        return hash(str(self.t) + str(self.q))

    def is_equal_except_time(self, state):
        return self.q == state.q

    def __str__(self):
        return str((self.t, self.q))


def create_intervals_from_tuples(tuple1, tuple2):
    return [Interval(lower=min(a, b), upper=max(a, b)) for a, b in zip(tuple1, tuple2)]
