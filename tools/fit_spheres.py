#!/usr/bin/env python
"""Offline sphere decomposition of the robot collision geometry.

Run in the build container only (it reads the reference's model files):

    python tools/fit_spheres.py            # writes mmmp_b200/geometry/*.json

Panda: every collision mesh models/franka_panda/meshes/collision/*.stl is
already a convex polytope (all vertices lie on their hull), which is also how
Bullet treats a URDF mesh.  Each hull is covered by K spheres (k-centre
clustering of a dense point sample of the hull, each cluster replaced by its
minimum enclosing ball + pad), so the union of spheres CONTAINS the hull
(no false "free" verdicts) and protrudes from it by at most `margin` (the
stated sphere-fit margin; false "collision" verdicts can only come from it).

KUKA iiwa14: models/kuka_iiwa/urdf/iiwa14_spheres_collision.urdf already is a
sphere model; it is parsed, not fitted.

The emitted JSON files are data derived from the meshes (centres/radii,
joint origins, joint limits), not code.
"""
from __future__ import annotations

import json
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial import ConvexHull

REF = os.environ.get("MMMP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mmmp_b200", "geometry")


def load_stl(path):
    b = open(path, "rb").read()
    n = struct.unpack("<I", b[80:84])[0]
    rec = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])
    a = np.frombuffer(b, dtype=rec, count=n, offset=84)
    return a["v"].reshape(-1, 3).astype(np.float64)


def hull_planes(verts):
    h = ConvexHull(verts)
    # scipy: equations [n, d] with n.x + d <= 0 inside
    return h, h.equations[:, :3].copy(), -h.equations[:, 3].copy()


def sample_hull(verts, spacing, interior=True):
    """Point sample of the polytope: hull vertices, area-proportional surface points,
    and (optionally) interior grid points."""
    h, n, d = hull_planes(verts)
    pts = [verts[h.vertices]]
    rng = np.random.default_rng(12345)
    for simplex in h.simplices:
        a, b, c = verts[simplex]
        area = 0.5 * np.linalg.norm(np.cross(b - a, c - a))
        m = int(np.ceil(area / (spacing * spacing) * 1.5)) + 1
        u, v = rng.random(m), rng.random(m)
        flip = u + v > 1
        u[flip], v[flip] = 1 - u[flip], 1 - v[flip]
        pts.append(a[None] + u[:, None] * (b - a)[None] + v[:, None] * (c - a)[None])
        # edges of the facet
        for p0, p1 in ((a, b), (b, c), (c, a)):
            k = int(np.ceil(np.linalg.norm(p1 - p0) / spacing)) + 1
            t = np.linspace(0, 1, k)[:, None]
            pts.append(p0[None] * (1 - t) + p1[None] * t)
    if interior:
        lo, hi = verts.min(0), verts.max(0)
        axes = [np.arange(lo[i], hi[i] + 2 * spacing, 2 * spacing) for i in range(3)]
        g = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, 3)
        pts.append(g[(g @ n.T - d[None, :] <= 1e-9).all(1)])
    return np.concatenate(pts, 0), n, d


def fit_link(verts, eps, floor_z=None, floor_eps=None, cand_spacing=0.006, samp_spacing=0.003, pad=0.0004):
    """Greedy set cover: candidate centres on a grid inside the hull, each with the
    largest radius that protrudes at most `eps` (and at most `floor_eps` below the
    plane z = floor_z, used for the base link that sits on the ground)."""
    lo, hi = verts.min(0), verts.max(0)
    cand_spacing = min(cand_spacing, float((hi - lo).min()) / 7.0)
    samp_spacing = min(samp_spacing, cand_spacing * 0.6)
    pts, n, d = sample_hull(verts, samp_spacing)
    axes = [np.arange(lo[i], hi[i] + cand_spacing, cand_spacing) for i in range(3)]
    g = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, 3)
    depth = (d[None, :] - g @ n.T).min(1)
    keep = depth >= 0
    g, depth = g[keep], depth[keep]
    r = depth + eps
    if floor_z is not None:
        r = np.minimum(r, g[:, 2] - floor_z + floor_eps)
    ok = r > 0.003
    g, r, depth = g[ok], r[ok], depth[ok]
    cov = np.zeros((len(g), len(pts)), dtype=bool)
    for i0 in range(0, len(g), 512):
        D2 = ((g[i0:i0 + 512, None, :] - pts[None, :, :]) ** 2).sum(-1)
        cov[i0:i0 + 512] = D2 <= (r[i0:i0 + 512, None] - pad) ** 2
    unc = np.ones(len(pts), bool)
    chosen = []
    while unc.any():
        gain = (cov & unc[None, :]).sum(1)
        j = int(np.argmax(gain))
        if gain[j] == 0:
            break
        chosen.append(j)
        unc &= ~cov[j]
    centres, radii = g[chosen].copy(), r[chosen].copy()
    # tighten: a sphere only needs to reach the points no other sphere covers
    D = np.sqrt(((centres[:, None, :] - pts[None]) ** 2).sum(-1))
    for _ in range(2):
        for j in range(len(centres)):
            others = np.delete(np.arange(len(centres)), j)
            covered_by_others = (D[others] <= radii[others, None] - pad).any(0)
            need = D[j][~covered_by_others]
            radii[j] = min(radii[j], (need.max() + pad) if len(need) else 0.0)
    keep = radii > 0
    centres, radii = centres[keep], radii[keep]
    depth = (d[None, :] - centres @ n.T).min(1)
    margin = float((radii - depth).max())
    chk, _, _ = sample_hull(verts, samp_spacing * 0.7)
    unc_d = (np.sqrt(((chk[:, None, :] - centres[None]) ** 2).sum(-1)) - radii[None]).min(1)
    # coverage pad: the cover was built on a point sample; inflate by the worst gap the finer
    # check sample shows (+0.3 mm) so that the union of spheres contains the hull for certain
    cov_pad = max(float(unc_d.max()), 0.0) + 0.0003
    radii = radii + cov_pad
    margin += cov_pad
    return centres, radii, margin, float(unc_d.max()) - cov_pad, int(unc.sum())


def rpy_matrix(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def fit_panda(ks):
    mesh_dir = os.path.join(REF, "models/franka_panda/meshes/collision")
    # link name, mesh, frame (0 = base, f = after joint f-1), link id, transform link->frame
    hand_R = rpy_matrix((0, 0, -0.785398163397))                  # panda_arm_hand.urdf:194
    T_hand = np.eye(4); T_hand[:3, :3] = hand_R; T_hand[2, 3] = 0.107   # joint8 :186 then hand joint
    T_lf = T_hand @ np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0.0584], [0, 0, 0, 1.0]])  # :237
    T_rf = T_lf @ np.array([[-1, 0, 0, 0], [0, -1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])        # :228 rpy z=pi
    parts = [("panda_link0", "link0", 0, 0, np.eye(4))]
    for i in range(1, 8):
        parts.append((f"panda_link{i}", f"link{i}", i, i, np.eye(4)))
    parts += [("panda_hand", "hand", 7, 8, T_hand), ("panda_leftfinger", "finger", 7, 9, T_lf),
              ("panda_rightfinger", "finger", 7, 10, T_rf)]
    spheres, report = [], {}
    for name, mesh, frame, link, T in parts:
        verts = load_stl(os.path.join(mesh_dir, mesh + ".stl"))
        if name == "panda_link0":  # bolted 1 cm above the ground plane in the reference's scenes
            c, r, margin, uncovered, left = fit_link(verts, ks[name], floor_z=0.0, floor_eps=0.008)
        else:
            c, r, margin, uncovered, left = fit_link(verts, ks[name])
        cw = c @ T[:3, :3].T + T[:3, 3]
        for ci, ri in zip(cw, r):
            spheres.append({"link": name, "link_id": link, "frame": frame,
                            "center": [float(x) for x in ci], "radius": float(ri)})
        report[name] = {"eps_target_m": ks[name], "n": int(len(r)), "margin_m": margin, "uncovered_m": uncovered,
                        "r_min": float(r.min()), "r_max": float(r.max())}
        print(f"{name:18s} n={len(r):3d} margin={margin*1000:6.2f} mm uncovered={uncovered*1000:6.3f} mm left={left} "
              f"r=[{r.min()*1000:.1f},{r.max()*1000:.1f}] mm", file=sys.stderr)
    return spheres, report


def panda_model(ks):
    spheres, report = fit_panda(ks)
    # modified DH rows robots/panda.py:31-40 ; joint limits panda_arm_hand.urdf:62..182
    dh = [(0, 0.333, 0), (0, 0, -np.pi / 2), (0, 0.316, np.pi / 2), (0.0825, 0, np.pi / 2),
          (-0.0825, 0.384, -np.pi / 2), (0, 0, np.pi / 2), (0.088, 0, np.pi / 2), (0, 0.107, 0)]
    limits = [(-2.8973, 2.8973), (-1.7628, 1.7628), (-2.8973, 2.8973), (-3.0718, -0.0698),
              (-2.8973, 2.8973), (-0.0175, 3.7525), (-2.8973, 2.8973)]
    return {
        "name": "franka_panda",
        "source": "models/franka_panda/urdf/panda_arm_hand.urdf + meshes/collision/*.stl (convex hulls)",
        "n_joints": 7,
        "dh_modified": [{"a": a, "d": d, "alpha": al} for a, d, al in dh],
        "joint_limits": limits,
        "links": ["panda_link0"] + [f"panda_link{i}" for i in range(1, 8)] +
                 ["panda_hand", "panda_leftfinger", "panda_rightfinger"],
        # pybullet link index i = panda_link(i+1); env.py:79-80 pairs i in 0..4, j in i+2..6
        "self_pairs": [[i + 1, j + 1] for i in range(5) for j in range(i + 2, 7)],
        "spheres": spheres,
        "fit_report": report,
        "max_margin_m": max(v["margin_m"] for v in report.values()),
    }


def panda_pair_table(n_u=256, n_v=384):
    """Exact decision table for the (panda_link5, panda_link7) hull pair: the relative pose
    frame5^-1 frame7 = O5 Rz(q5) O6 Rz(q6) depends on joints 5 and 6 (0-based) only.
    Cell (iu, iv) is marked when the hull distance at any of its corners is within the
    cell's Lipschitz bound, so a clear bit PROVES the pair is collision free in the cell."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import hull as H  # offline tool: allowed to use the checker
    from oracle.chain import PANDA_DH, mdh
    hulls = H.hulls()
    v5, v7 = hulls[5], hulls[7]
    lim = [(-0.0175, 3.7525), (-2.8973, 2.8973)]
    u = np.linspace(lim[0][0], lim[0][1], n_u + 1)
    v = np.linspace(lim[1][0], lim[1][1], n_v + 1)
    U, V = np.meshgrid(u, v, indexing="ij")
    a5, d5, al5 = PANDA_DH[5]
    a6, d6, al6 = PANDA_DH[6]
    T = np.matmul(mdh(a5, d5, al5, U.ravel()), mdh(a6, d6, al6, V.ravel()))
    eye = np.broadcast_to(np.eye(4), T.shape).copy()
    dist = H.posed_distance(v5, eye, v7, T).reshape(n_u + 1, n_v + 1)
    # Lipschitz bound: points of link7 move at most R_v per rad of q6 and R_u per rad of q5
    R_v = float(np.sqrt(v7[:, 0] ** 2 + v7[:, 1] ** 2).max())
    R_u = float(abs(a6) + np.linalg.norm(v7, axis=1).max())
    du, dv = (lim[0][1] - lim[0][0]) / n_u, (lim[1][1] - lim[1][0]) / n_v
    delta = R_u * du / 2 + R_v * dv / 2
    c = np.minimum(np.minimum(dist[:-1, :-1], dist[1:, :-1]), np.minimum(dist[:-1, 1:], dist[1:, 1:]))
    bits = (c <= delta)
    flat = bits.ravel().astype(np.uint8)
    words = np.zeros((len(flat) + 31) // 32, dtype=np.uint32)
    idx = np.nonzero(flat)[0]
    np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    exact = float((dist <= 0).mean())
    print(f"pair table link5/link7: {n_u}x{n_v}, lipschitz delta={delta*1000:.2f} mm, "
          f"marked {bits.mean():.4f} of cells, exact colliding corners {exact:.4f}", file=sys.stderr)
    return {"link_a": 5, "link_b": 7, "joint_u": 5, "joint_v": 6, "n_u": n_u, "n_v": n_v,
            "lo_u": lim[0][0], "hi_u": lim[0][1], "lo_v": lim[1][0], "hi_v": lim[1][1],
            "lipschitz_delta_m": delta, "words_hex": words.tobytes().hex()}


def kuka_model():
    path = os.path.join(REF, "models/kuka_iiwa/urdf/iiwa14_spheres_collision.urdf")
    root = ET.parse(path).getroot()
    links = {}
    for link in root.findall("link"):
        cols = []
        for col in link.findall("collision"):
            org = col.find("origin")
            xyz = [float(x) for x in org.get("xyz", "0 0 0").split()] if org is not None else [0, 0, 0]
            geom = col.find("geometry")
            sp = geom.find("sphere")
            cy = geom.find("cylinder")
            if sp is not None:
                cols.append({"center": xyz, "radius": float(sp.get("radius"))})
            elif cy is not None:
                # base cylinder: cover with spheres stacked along z
                r, L = float(cy.get("radius")), float(cy.get("length"))
                nz = max(1, int(np.ceil(L / r)))
                for z in np.linspace(-L / 2 + L / (2 * nz), L / 2 - L / (2 * nz), nz):
                    cols.append({"center": [xyz[0], xyz[1], xyz[2] + float(z)],
                                 "radius": float(np.hypot(r, L / (2 * nz)))})
        links[link.get("name")] = cols
    joints = []
    for j in root.findall("joint"):
        if j.get("type") != "revolute":
            continue
        org = j.find("origin")
        lim = j.find("limit")
        joints.append({
            "name": j.get("name"), "parent": j.find("parent").get("link"), "child": j.find("child").get("link"),
            "xyz": [float(x) for x in org.get("xyz").split()], "rpy": [float(x) for x in org.get("rpy").split()],
            "lower": float(lim.get("lower")), "upper": float(lim.get("upper")),
        })
    order = [joints[0]["parent"]] + [j["child"] for j in joints]
    spheres = []
    for f, name in enumerate(order):
        for c in links.get(name, []):
            spheres.append({"link": name, "link_id": f, "frame": f, "center": c["center"], "radius": c["radius"]})
    return {
        "name": "kuka_iiwa14",
        "source": "models/kuka_iiwa/urdf/iiwa14_spheres_collision.urdf",
        "n_joints": len(joints),
        "joint_origins": [{"xyz": j["xyz"], "rpy": j["rpy"]} for j in joints],
        "tip": {"xyz": [0, 0, 0.045], "rpy": [0, 0, 0]},
        "joint_limits": [(j["lower"], j["upper"]) for j in joints],
        "links": order,
        # env.py:79-80: pybullet link indices i in 0..4, j in i+2..6 = links 1..7; the base link is never self-checked
        "self_pairs": [[i, j] for i in range(1, len(order)) for j in range(i + 2, len(order))],
        "spheres": spheres,
        "max_margin_m": 0.0,
    }


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    # per-link protrusion targets [m]
    ks = {"panda_link0": 0.020, "panda_link1": 0.015, "panda_link2": 0.015, "panda_link3": 0.015,
          "panda_link4": 0.015, "panda_link5": 0.012, "panda_link6": 0.012, "panda_link7": 0.010,
          "panda_hand": 0.010, "panda_leftfinger": 0.006, "panda_rightfinger": 0.006}
    if len(sys.argv) > 1:
        ks.update(json.loads(sys.argv[1]))
    m = panda_model(ks)
    m["pair_tables"] = [panda_pair_table()]
    json.dump(m, open(os.path.join(OUT, "panda_spheres.json"), "w"), indent=1)
    print("panda: %d spheres, max margin %.1f mm" % (len(m["spheres"]), m["max_margin_m"] * 1000), file=sys.stderr)
    k = kuka_model()
    json.dump(k, open(os.path.join(OUT, "kuka_iiwa14_spheres.json"), "w"), indent=1)
    print("kuka: %d spheres, %d joints" % (len(k["spheres"]), k["n_joints"]), file=sys.stderr)
