"""Sphere-model collision checker in float64 numpy (TEST INFRASTRUCTURE).

Restates the collision MODEL the CUDA kernels implement (mmmp_b200/csrc/collide.cu):
every collision link is a union of spheres fitted offline from the URDF hulls
(mmmp_b200/geometry/*.json), tested against obstacle primitives, the robot's own
link pairs of planners/prm/pybullet/env.py:79-80 and other robots' links
(env.py:51-84), touching = collision.  It checks that the kernels implement the
model; oracle/hull.py measures how far the model is from the reference's hulls.
"""
import numpy as np

from . import chain


def model_origins(model):
    if "dh_modified" in model:
        out = []
        for r in model["dh_modified"]:
            a, d, al = r["a"], r["d"], r["alpha"]
            ca, sa = np.cos(al), np.sin(al)
            out.append(np.array([[1.0, 0, 0, a], [0, ca, -sa, -sa * d], [0, sa, ca, ca * d], [0, 0, 0, 1.0]]))
        return out
    org = [chain.origin_matrix(j["xyz"], j["rpy"]) for j in model["joint_origins"]]
    org.append(chain.origin_matrix(model["tip"]["xyz"], model["tip"]["rpy"]))
    return org


def frames(model, q, base):
    """[N, n+1, 4, 4]: frame 0 = base, frame f = after joint f-1."""
    q = np.atleast_2d(np.asarray(q, np.float64))
    fr = chain.chain_frames(model_origins(model), q, base)      # [N, n+1] : frames 1..n and tip
    N = q.shape[0]
    T0 = np.broadcast_to(base, (N, 1, 4, 4))
    return np.concatenate([T0, fr[:, :-1]], 1)


def world_spheres(model, q, base):
    """-> centres [N, S, 3], radii [S], link ids [S]."""
    F = frames(model, q, base)
    sph = sorted(model["spheres"], key=lambda s: s["link_id"])
    c = np.array([s["center"] for s in sph])
    r = np.array([s["radius"] for s in sph])
    f = np.array([s["frame"] for s in sph])
    l = np.array([s["link_id"] for s in sph])
    T = F[:, f]                                                   # [N, S, 4, 4]
    w = np.einsum("nsij,sj->nsi", T[..., :3, :3], c) + T[..., :3, 3]
    return w, r, l


def sphere_obstacle_hit(w, r, ob):
    """w [..., 3] centres, r radii broadcastable -> bool [...]."""
    t = ob["type"]
    if t == "plane":
        return w @ np.asarray(ob["normal"]) - ob["offset"] <= r
    if t == "sphere":
        return np.linalg.norm(w - np.asarray(ob["center"]), axis=-1) <= r + ob["radius"]
    if t == "box":
        R = np.asarray(ob["R"], float)
        b = (w - np.asarray(ob["center"])) @ R
        e = np.maximum(np.abs(b) - np.asarray(ob["half"]), 0.0)
        return (e ** 2).sum(-1) <= r ** 2
    if t == "cylinder":
        d = w - np.asarray(ob["center"])
        h = d @ np.asarray(ob["axis"])
        rad = np.sqrt(np.maximum((d ** 2).sum(-1) - h ** 2, 0.0))
        eh = np.maximum(np.abs(h) - ob["half_len"], 0.0)
        er = np.maximum(rad - ob["radius"], 0.0)
        return eh ** 2 + er ** 2 <= r ** 2
    raise ValueError(t)


def table_hit(model, q):
    """Exact pair tables (bit lookup) -> bool [N]."""
    q = np.atleast_2d(np.asarray(q, np.float64))
    hit = np.zeros(q.shape[0], bool)
    for pt in model.get("pair_tables", []):
        words = np.frombuffer(bytes.fromhex(pt["words_hex"]), dtype=np.uint32)
        qu = q[:, pt["joint_u"]].astype(np.float32)
        qv = q[:, pt["joint_v"]].astype(np.float32)
        inv_du = np.float32(pt["n_u"] / (pt["hi_u"] - pt["lo_u"]))
        inv_dv = np.float32(pt["n_v"] / (pt["hi_v"] - pt["lo_v"]))
        iu = np.clip(np.floor((qu - np.float32(pt["lo_u"])) * inv_du).astype(np.int64), 0, pt["n_u"] - 1)
        iv = np.clip(np.floor((qv - np.float32(pt["lo_v"])) * inv_dv).astype(np.int64), 0, pt["n_v"] - 1)
        cell = iu * pt["n_v"] + iv
        hit |= ((words[cell >> 5] >> (cell & 31).astype(np.uint32)) & 1).astype(bool)
    return hit


def self_hit(model, w, r, l, q=None, use_tables=True):
    pairs = {(a, b) for a, b in model["self_pairs"]}
    hit = np.zeros(w.shape[0], bool)
    if use_tables:
        for a, b in model.get("proven_free_pairs", []):      # proven never to collide: not tested
            pairs.discard((a, b))
            pairs.discard((b, a))
    if use_tables and q is not None:
        hit |= table_hit(model, q)
        for pt in model.get("pair_tables", []):
            pairs.discard((pt["link_a"], pt["link_b"]))
            pairs.discard((pt["link_b"], pt["link_a"]))
    for a, b in pairs:
        ia, ib = np.nonzero(l == a)[0], np.nonzero(l == b)[0]
        if len(ia) == 0 or len(ib) == 0:
            continue
        d = np.linalg.norm(w[:, ia, None, :] - w[:, None, ib, :], axis=-1)
        hit |= (d <= (r[ia][:, None] + r[ib][None, :])).any((1, 2))
    return hit


class SphereChecker:
    def __init__(self, models, bases, obstacles=(), use_tables=True):
        self.models, self.bases, self.obstacles = list(models), [np.asarray(b, float) for b in bases], list(obstacles)
        self.use_tables = use_tables

    def config_collision(self, q, r=0, self_=True, obstacles=True):
        w, rad, l = world_spheres(self.models[r], q, self.bases[r])
        hit = np.zeros(w.shape[0], bool)
        if obstacles:
            for ob in self.obstacles:
                hit |= sphere_obstacle_hit(w, rad[None, :], ob).any(1)
        if self_:
            hit |= self_hit(self.models[r], w, rad, l, q, self.use_tables)
        return hit

    def robot_robot_collision(self, q1, q2, r1, r2):
        w1, ra, l1 = world_spheres(self.models[r1], q1, self.bases[r1])
        w2, rb, l2 = world_spheres(self.models[r2], q2, self.bases[r2])
        hit = np.zeros(w1.shape[0], bool)
        # base-base pairs are configuration independent and skipped by the kernels
        f1 = np.array([s["frame"] for s in sorted(self.models[r1]["spheres"], key=lambda s: s["link_id"])])
        f2 = np.array([s["frame"] for s in sorted(self.models[r2]["spheres"], key=lambda s: s["link_id"])])
        for n0 in range(0, w1.shape[0], 256):
            d = np.linalg.norm(w1[n0:n0 + 256, :, None, :] - w2[n0:n0 + 256, None, :, :], axis=-1)
            ok = d <= (ra[:, None] + rb[None, :])
            ok &= ~((f1[:, None] == 0) & (f2[None, :] == 0))
            hit[n0:n0 + 256] = ok.any((1, 2))
        return hit

    def composite_collision(self, q, self_=True, obstacles=True, robots=True):
        """CoupledPRM.is_collision_free_sample negated (coupled_prm.py:143-160); q [N, sum D]."""
        q = np.atleast_2d(q)
        hit = np.zeros(q.shape[0], bool)
        off = [0]
        for m in self.models:
            off.append(off[-1] + m["n_joints"])
        for r in range(len(self.models)):
            hit |= self.config_collision(q[:, off[r]:off[r + 1]], r, self_, obstacles)
        if robots:
            for a in range(len(self.models)):
                for b in range(a + 1, len(self.models)):
                    hit |= self.robot_robot_collision(q[:, off[a]:off[a + 1]], q[:, off[b]:off[b + 1]], a, b)
        return hit

    def edge_collision(self, qa, qb, n_steps, step, r=0, **kw):
        """is_collision_free_edge negated (decoupled_prm.py:855-863): OR over t = i*step."""
        qa, qb = np.atleast_2d(qa), np.atleast_2d(qb)
        hit = np.zeros(qa.shape[0], bool)
        for i in range(n_steps):
            t = i * step
            q = qa + t * (qb - qa)
            hit |= self.config_collision(q, r, **kw) if r >= 0 else self.composite_collision(q, **kw)
        return hit
