"""Planar arms: FK and analytic collision predicates, float64 scalar Python in the
reference's operation order (TEST INFRASTRUCTURE).

Follows robots/planar_arm.py:23-40 (forward_kinematics) and
planners/prm/planar/env.py:78-229 (robot_robot_collision, joint_link_collision,
link_link_collision, line_intersects_rect/circle, robot_self_collision,
robot_in_env_bounds) plus the wrappers single_arm/prm.py:104-169,
multi_arm/coupled/coupled_prm.py:83-109, multi_arm/decoupled/decoupled_prm.py:126-152.
Pinned against the reference itself (tools/gen_golden.py -> tests/golden/planar_*.npz).
"""
import math

import numpy as np


def fk(links, base, q):
    """planar_arm.py:23-40 -> list of (x, y) link end points."""
    pos = []
    x, y = float(base[0]), float(base[1])
    cum = 0.0
    for i, L in enumerate(links):
        cum = float(q[0]) if i == 0 else cum + float(q[i])
        x = x + L * math.cos(cum)
        y = y + L * math.sin(cum)
        pos.append((x, y))
    return pos


def ccw(a, b, c):
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]) > 0


def seg_seg(p1, q1, p2, q2):  # env.py:132-141
    return ccw(p1, q1, p2) != ccw(p1, q1, q2) and ccw(p2, q2, p1) != ccw(p2, q2, q1)


def joint_link(j, lo, hi, radius=0.5):  # env.py:119-130
    lx, ly = hi[0] - lo[0], hi[1] - lo[1]
    jx, jy = j[0] - lo[0], j[1] - lo[1]
    den = lx * lx + ly * ly
    if den == 0.0:
        return False
    proj = (jx * lx + jy * ly) / den
    if proj < 0 or proj > 1:
        return False
    return abs(lx * jy - ly * jx) / math.sqrt(den) <= radius


def seg_rect(p, q, o):  # env.py:168-188 ; o = (x0, y0, x1, y1)
    v = [(o[0], o[1]), (o[2], o[1]), (o[2], o[3]), (o[0], o[3])]
    return any(seg_seg(p, q, v[i], v[(i + 1) % 4]) for i in range(4))


def seg_circle(p, q, o):  # env.py:190-205 ; o = (cx, cy, r)
    vx, vy = o[0] - p[0], o[1] - p[1]
    dx, dy = q[0] - p[0], q[1] - p[1]
    den = dx * dx + dy * dy
    if den == 0.0:
        return False
    t = (vx * dx + vy * dy) / den
    cx, cy = p[0] + t * dx, p[1] + t * dy
    return math.sqrt((cx - o[0]) ** 2 + (cy - o[1]) ** 2) <= o[2] and 0 <= t <= 1


def in_bounds(b, p):
    return b[0] <= p[0] <= b[1] and b[2] <= p[1] <= b[3]


def config_in_collision(world, q, arm=-1, bounds=True, self_=True, obstacles=True, robots=True):
    """world = {arms: [{links, base}], bounds: (xmin,xmax,ymin,ymax), obstacles: [...]}.
    arm >= 0: q holds that arm's joints; arm < 0: composite row."""
    arms = world["arms"]
    sel = range(len(arms)) if arm < 0 else [arm]
    pos = {}
    off = 0
    for a in range(len(arms)):
        n = len(arms[a]["links"])
        if a in sel:
            qa = q[off:off + n] if arm < 0 else q[:n]
            pos[a] = fk(arms[a]["links"], arms[a]["base"], qa)
        off += n
    for a in sel:
        base = tuple(arms[a]["base"])
        P = pos[a]
        n = len(P)
        if bounds:
            if not in_bounds(world["bounds"], base) or any(not in_bounds(world["bounds"], p) for p in P):
                return True
        if self_:
            for i in range(n):
                prev = base if i == 0 else P[i - 1]
                for j in range(i + 2, n):
                    if seg_seg(P[i], prev, P[j], P[j - 1]):
                        return True
        if obstacles:
            for ob in world["obstacles"]:
                for i in range(n):
                    prev = base if i == 0 else P[i - 1]
                    if ob["type"] == "rect":
                        if seg_rect(P[i], prev, (ob["x0"], ob["y0"], ob["x1"], ob["y1"])):
                            return True
                    elif seg_circle(P[i], prev, (ob["cx"], ob["cy"], ob["r"])):
                        return True
    if robots and arm < 0:
        R = world.get("joint_link_radius", 0.5)
        for a in range(len(arms)):
            for b in range(a + 1, len(arms)):
                Pa, Pb = pos[a], pos[b]
                ba, bb = tuple(arms[a]["base"]), tuple(arms[b]["base"])
                for i in range(len(Pa)):
                    prev = ba if i == 0 else Pa[i - 1]
                    for j in range(len(Pb)):
                        prev2 = bb if j == 0 else Pb[j - 1]
                        if seg_seg(Pa[i], prev, Pb[j], prev2):
                            return True
                for i in range(len(Pa)):
                    for j in range(len(Pb)):
                        lo = bb if j == 0 else Pb[j - 1]
                        if joint_link(Pa[i], lo, Pb[j], R):
                            return True
                for i in range(len(Pb)):
                    for j in range(len(Pa)):
                        lo = ba if j == 0 else Pa[j - 1]
                        if joint_link(Pb[i], lo, Pa[j], R):
                            return True
    return False


def edge_in_collision(world, q1, q2, local_step, arm=-1, **kw):
    """is_collision_free_edge negated, single_arm/prm.py:161-169."""
    q1, q2 = np.asarray(q1, float), np.asarray(q2, float)
    for t in np.arange(0, 1, local_step):
        if config_in_collision(world, q1 + t * (q2 - q1), arm, **kw):
            return True
    return False
