#!/usr/bin/env python
"""Offline analysis of the Panda self-collision link pairs with the hull checker
(oracle/hull_gjk.c; build container or GPU box, no reference tree needed).

The relative pose of links a < b depends only on the joints between them.  For every pair
of planners/prm/pybullet/env.py:79-80 whose relative pose depends on 2 or 3 joints, the
hull distance is evaluated on a regular grid over those joints' ranges; with the
Lipschitz bound delta = sum_j R_j * dq_j / 2 (R_j = largest distance of a point of link b
from joint j's axis) a pair whose grid minimum exceeds delta is PROVEN collision free over
the whole joint range and is dropped from the self-collision mask; a 2-joint pair that does
collide gets an exact bit table (cell marked when any corner is within delta).

Updates mmmp_b200/geometry/panda_spheres.json in place: "proven_free_pairs", "pair_tables".
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import hull as H  # noqa: E402
from oracle.chain import PANDA_DH, mdh  # noqa: E402

PATH = os.path.join(ROOT, "mmmp_b200", "geometry", "panda_spheres.json")


def rel_pose(a, b, qs):
    """frame_a^-1 frame_b for joint values qs[:, j], j = a .. b-1 (frame f = after joint f-1)."""
    T = np.broadcast_to(np.eye(4), (qs.shape[0], 4, 4)).copy()
    for c, j in enumerate(range(a, b)):
        al, d, alpha = PANDA_DH[j]
        T = np.matmul(T, mdh(al, d, alpha, qs[:, c]))
    return T


def radii(a, b, vb):
    """R_j for j = a..b-1: bound on the distance of link b's points from joint j's axis."""
    out = []
    for j in range(a, b):
        # points of link b expressed in frame j+1 for all poses: bound by chain offsets + |p|
        reach = float(np.linalg.norm(vb, axis=1).max())
        for i in range(j + 1, b):
            al, d, _ = PANDA_DH[i]
            reach += float(np.hypot(al, d))
        out.append(reach)
    return out


def analyse(model, a, b, cells):
    hulls = H.hulls()
    lim = np.array(model["joint_limits"])[a:b]
    va, vb = hulls[a], hulls[b]
    axes = [np.linspace(lim[c, 0], lim[c, 1], cells[c] + 1) for c in range(b - a)]
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, b - a)
    T = rel_pose(a, b, grid)
    eye = np.broadcast_to(np.eye(4), T.shape).copy()
    dist = H.posed_distance(va, eye, vb, T).reshape([c + 1 for c in cells])
    R = radii(a, b, vb)
    delta = sum(R[c] * (lim[c, 1] - lim[c, 0]) / cells[c] / 2 for c in range(b - a))
    return dist, delta, lim


def pack_bits(bits):
    flat = bits.ravel().astype(np.uint8)
    words = np.zeros((len(flat) + 31) // 32, dtype=np.uint32)
    idx = np.nonzero(flat)[0]
    np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    return words


if __name__ == "__main__":
    model = json.load(open(PATH))
    proven, tables = [], []
    for a, b in model["self_pairs"]:
        m = b - a
        if m == 2:
            cells = [256, 384]
        elif m == 3:
            cells = [72, 72, 72]
        else:
            continue
        dist, delta, lim = analyse(model, a, b, cells)
        dmin = float(dist.min())
        status = "PROVEN FREE" if dmin > delta else ("table" if m == 2 else "kept (spheres)")
        print(f"pair ({a},{b}): joints {a}..{b - 1}, grid {cells}, min hull distance {dmin * 1000:7.1f} mm, "
              f"lipschitz delta {delta * 1000:6.1f} mm -> {status}", file=sys.stderr)
        if dmin > delta:
            proven.append([a, b])
        elif m == 2:
            c = dist
            corner_min = np.minimum(np.minimum(c[:-1, :-1], c[1:, :-1]), np.minimum(c[:-1, 1:], c[1:, 1:]))
            bits = corner_min <= delta
            tables.append({"link_a": a, "link_b": b, "joint_u": a, "joint_v": a + 1, "n_u": cells[0], "n_v": cells[1],
                           "lo_u": float(lim[0, 0]), "hi_u": float(lim[0, 1]), "lo_v": float(lim[1, 0]),
                           "hi_v": float(lim[1, 1]), "lipschitz_delta_m": delta,
                           "marked_fraction": float(bits.mean()), "words_hex": pack_bits(bits).tobytes().hex()})
    model["proven_free_pairs"] = proven
    model["pair_tables"] = tables
    json.dump(model, open(PATH, "w"), indent=1)
    print(f"proven free: {proven}; tables: {[(t['link_a'], t['link_b']) for t in tables]}", file=sys.stderr)
