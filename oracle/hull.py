"""Convex-hull collision oracle for the Panda (TEST INFRASTRUCTURE; PARITY UNPINNED).

Restates what the reference asks pybullet for (planners/prm/pybullet/env.py:51-84):
  robot_self_collision : link pairs (i, j), i in 0..4, j in i+2..6 (pybullet link index
                         = panda_link(i+1)), closest distance <= 0          env.py:77-84
  robot_obstacle_collision : whole robot (11 bodies) vs one obstacle body    env.py:51-68
  robot_robot_collision    : whole robot vs whole robot                      env.py:70-75
Geometry = convex hulls of the URDF collision meshes (oracle/data/panda_hulls.npz,
exported by tools/export_hulls.py) posed by the float64 chain of oracle/chain.py.
`margin` is added per body (Bullet's collision margin; 0 by default).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import chain

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LINKS = ["panda_link0"] + [f"panda_link{i}" for i in range(1, 8)] + ["panda_hand", "panda_leftfinger", "panda_rightfinger"]
# frame index (0 = base, f = DH frame f-1) every link's vertices are expressed in
LINK_FRAME = [0, 1, 2, 3, 4, 5, 6, 7, 7, 7, 7]
SELF_PAIRS = [(i + 1, j + 1) for i in range(5) for j in range(i + 2, 7)]  # env.py:79-80 (link ids)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "libhull_gjk.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(so)
        L.hull_gjk_distance.restype = C.c_double
        L.hull_gjk_distance.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_double]
        L.hull_gjk_posed_batch.restype = None
        L.hull_gjk_posed_batch.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hull_plane_posed_batch.restype = None
        L.hull_plane_posed_batch.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                                             C.c_void_p]
        _LIB = L
    return _LIB


_HULLS = None


def hulls():
    global _HULLS
    if _HULLS is None:
        z = np.load(os.path.join(_HERE, "data", "panda_hulls.npz"))
        _HULLS = [np.ascontiguousarray(z[name], dtype=np.float64) for name in LINKS]
    return _HULLS


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def distance(va, vb, ra=0.0, rb=0.0):
    va = np.ascontiguousarray(va, dtype=np.float64)
    vb = np.ascontiguousarray(vb, dtype=np.float64)
    return lib().hull_gjk_distance(_ptr(va), len(va), ra, _ptr(vb), len(vb), rb)


def posed_distance(va, TA, vb, TB, ra=0.0, rb=0.0):
    """va under poses TA [n,3,4] vs vb under TB [n,3,4] -> [n] distances."""
    TA = np.ascontiguousarray(TA[:, :3, :], dtype=np.float64)
    TB = np.ascontiguousarray(TB[:, :3, :], dtype=np.float64)
    n = TA.shape[0]
    out = np.empty(n)
    scratch = np.empty((len(va) + len(vb)) * 3)
    lib().hull_gjk_posed_batch(_ptr(va), len(va), ra, _ptr(TA), _ptr(vb), len(vb), rb, _ptr(TB), n, _ptr(scratch),
                               _ptr(out))
    return out


def link_poses(q, base=None):
    """[N, 8 (base + 7 DH frames 0..6), 4, 4] poses the 11 link hulls hang off."""
    fr = chain.panda_frames(q, base)
    N = fr.shape[0]
    T0 = np.broadcast_to(np.eye(4) if base is None else base, (N, 1, 4, 4))
    return np.concatenate([T0, fr[:, :7]], 1)


def obstacle_vertices(ob):
    """World-space vertex set (+ radius) of a primitive dict {type, ...} (same dicts as
    mmmp_b200.planners.prm.pybullet.env obstacles).  Cylinders become 96-gon prisms."""
    t = ob["type"]
    if t == "box":
        h = np.asarray(ob["half"], float)
        R = np.asarray(ob.get("R", np.eye(3)), float)
        c = np.asarray(ob["center"], float)
        s = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)
        return (s * h) @ R.T + c, 0.0
    if t == "sphere":
        return np.asarray([ob["center"]], float), float(ob["radius"])
    if t == "cylinder":
        c = np.asarray(ob["center"], float)
        ax = np.asarray(ob["axis"], float)
        ax = ax / np.linalg.norm(ax)
        u = np.cross(ax, [1.0, 0, 0]) if abs(ax[0]) < 0.9 else np.cross(ax, [0, 1.0, 0])
        u /= np.linalg.norm(u)
        v = np.cross(ax, u)
        th = np.linspace(0, 2 * np.pi, 96, endpoint=False)
        ring = ob["radius"] * (np.cos(th)[:, None] * u + np.sin(th)[:, None] * v)
        return np.concatenate([c + ob["half_len"] * ax + ring, c - ob["half_len"] * ax + ring]), 0.0
    raise ValueError(t)


class PandaHullChecker:
    """Float64 hull checker for R Pandas + primitive obstacles."""

    def __init__(self, bases, obstacles=(), margin=0.0):
        self.bases = [np.asarray(b, float) for b in bases]
        self.obstacles = list(obstacles)
        self.margin = float(margin)

    def self_collision(self, q, r=0):
        P = link_poses(q, self.bases[r])
        H = hulls()
        hit = np.zeros(P.shape[0], bool)
        for a, b in SELF_PAIRS:
            d = posed_distance(H[a], P[:, LINK_FRAME[a]], H[b], P[:, LINK_FRAME[b]], self.margin, self.margin)
            hit |= d <= 0.0
        return hit

    def obstacle_collision(self, q, r=0):
        P = link_poses(q, self.bases[r])
        H = hulls()
        N = P.shape[0]
        hit = np.zeros(N, bool)
        eye = np.broadcast_to(np.eye(4), (N, 4, 4))
        for ob in self.obstacles:
            for l in range(len(LINKS)):
                if ob["type"] == "plane":
                    out = np.empty(N)
                    TA = np.ascontiguousarray(P[:, LINK_FRAME[l], :3, :])
                    pl = np.asarray(list(ob["normal"]) + [ob["offset"]], float)
                    lib().hull_plane_posed_batch(_ptr(H[l]), len(H[l]), self.margin, _ptr(TA), _ptr(pl), N, _ptr(out))
                    hit |= out <= 0.0
                else:
                    vb, rb = obstacle_vertices(ob)
                    vb = np.ascontiguousarray(vb)
                    d = posed_distance(H[l], P[:, LINK_FRAME[l]], vb, eye, self.margin, rb)
                    hit |= d <= 0.0
        return hit

    def robot_robot_collision(self, q1, q2, r1, r2):
        P1, P2 = link_poses(q1, self.bases[r1]), link_poses(q2, self.bases[r2])
        H = hulls()
        hit = np.zeros(P1.shape[0], bool)
        for a in range(len(LINKS)):
            for b in range(len(LINKS)):
                d = posed_distance(H[a], P1[:, LINK_FRAME[a]], H[b], P2[:, LINK_FRAME[b]], self.margin, self.margin)
                hit |= d <= 0.0
        return hit

    def config_collision(self, q, r=0):
        """DecoupledPRM.is_collision_free_sample negated (decoupled_prm.py:840-850)."""
        return self.self_collision(q, r) | self.obstacle_collision(q, r)
