"""Neighbour search restatements (TEST INFRASTRUCTURE).

heap_knn      : the bounded-heap loop of connect_nearest_neighbors_degree,
                planners/prm/pybullet/decoupled/decoupled_prm.py:484-494
                (heap on (-dist, node); Node.__lt__ by id, utils.py:19-20).
greedy_degree : the accept pass decoupled_prm.py:496-503.
distance_lists: connect_nearest_neighbors_distance candidate lists :556-563.
brute_knn     : vectorised float64 ranking by (dist, id) for larger N.
shared_verdict_table : the candidate-table verdicts of mmmp_collide_knn_edges restated -- a pair of nodes
                that name each other is checked ONCE, from the lower id to the higher (the first of the
                two is_collision_free_edge calls the accept pass :496-503 can make for it: nodes are
                processed in id order); every other slot node -> candidate.
"""
import heapq

import numpy as np


class _N:
    __slots__ = ("id",)

    def __init__(self, i):
        self.id = i

    def __lt__(self, other):
        return self.id < other.id


def heap_knn(dist_fn, n, k1):
    """For every node i: list of (dist, id) of its k1 candidates, ascending (dist, id)."""
    nodes = [_N(i) for i in range(n)]
    out = []
    for i in range(n):
        h = []
        for j in range(n):
            if i == j:
                continue
            heapq.heappush(h, (-dist_fn(i, j), nodes[j]))
            if len(h) > k1:
                heapq.heappop(h)
        c = [(-d, nd) for d, nd in h]
        hh = []
        for v in c:
            heapq.heappush(hh, v)
        c = [heapq.heappop(hh) for _ in range(len(hh))]
        out.append([(d, nd.id) for d, nd in c])
    return out


def greedy_degree(cands, edge_free, k2, cap_other_end=True):
    """cands[i] = [(dist, id)...]; edge_free(i, j) -> bool.  Returns adjacency lists."""
    n = len(cands)
    adj = [[] for _ in range(n)]
    for i in range(n):
        for _, j in cands[i]:
            if len(adj[i]) == k2:
                break
            if edge_free(i, j) and j not in adj[i] and not (cap_other_end and len(adj[j]) == k2):
                adj[i].append(j)
                adj[j].append(i)
    return adj


def brute_knn(D, k, exclude_self=True):
    """D [nq, nr] float64 distance matrix -> (idx [nq,k], dist) by ascending (dist, id);
    exact float64 ties at the k boundary keep the LARGER ids (heap semantics)."""
    nq, nr = D.shape
    idx = np.empty((nq, k), dtype=np.int64)
    dist = np.empty((nq, k))
    for i in range(nq):
        d = D[i].copy()
        if exclude_self:
            d[i] = np.inf
        order = np.lexsort((np.arange(nr), d))
        sel = list(order[:k])
        if k < nr and d[order[k]] == d[order[k - 1]]:
            tie = d[order[k - 1]]
            lo = k - 1
            while lo > 0 and d[order[lo - 1]] == tie:
                lo -= 1
            hi = k
            while hi + 1 < nr and d[order[hi + 1]] == tie:
                hi += 1
            group = list(order[lo:hi + 1])
            need = k - lo
            sel = list(order[:lo]) + group[len(group) - need:]
        idx[i] = sel
        dist[i] = d[sel]
    return idx, dist


def shared_verdict_table(idx, check):
    """idx [n, k] candidate ids (-1 / out of range = empty slot); check(a, b) -> bool = is_collision_free_edge
    walked from a to b (decoupled_prm.py:855-863).  -> (free [n, k] bool, calls): the verdict of every slot and
    the list of directed checks actually made, one per distinct pair."""
    idx = np.asarray(idx)
    n, k = idx.shape
    names = [set(int(j) for j in row if 0 <= j < n) for row in idx]
    memo, calls = {}, []
    free = np.zeros((n, k), bool)
    for i in range(n):
        for s in range(k):
            j = int(idx[i, s])
            if j < 0 or j >= n:
                continue
            mutual = j != i and i in names[j]
            a, b = (min(i, j), max(i, j)) if mutual else (i, j)
            if (a, b) not in memo:
                memo[(a, b)] = bool(check(a, b))
                calls.append((a, b))
            free[i, s] = memo[(a, b)]
    return free, calls
