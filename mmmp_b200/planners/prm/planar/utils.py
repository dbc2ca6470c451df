"""planners/prm/planar/utils.py:1-14 of the reference."""
INF = float("inf")


class Node:
    __slots__ = ("id", "q")

    def __init__(self, id, q):
        self.id = id
        self.q = tuple([round(c, 2) for c in q])

    def __eq__(self, other):
        return self.id == other.id

    def __hash__(self):
        return hash(str(self.id) + str(self.q))

    def __str__(self):
        return str((self.id, self.q))
