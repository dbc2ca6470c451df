"""Serial-chain forward kinematics in float64 (TEST INFRASTRUCTURE).

Follows robots/panda.py:111-173 (modified-DH transform, forward_kinematics,
positions_from_fk) of the reference, batched over configurations.
"""
from __future__ import annotations

import numpy as np

# robots/panda.py:31-40
PANDA_DH = [
    (0.0, 0.333, 0.0), (0.0, 0.0, -np.pi / 2), (0.0, 0.316, np.pi / 2), (0.0825, 0.0, np.pi / 2),
    (-0.0825, 0.384, -np.pi / 2), (0.0, 0.0, np.pi / 2), (0.088, 0.0, np.pi / 2), (0.0, 0.107, 0.0),
]


def quat_to_matrix(q):
    """xyzw quaternion -> rotation matrix (scipy Rotation.from_quat(...).as_matrix(), panda.py:127)."""
    x, y, z, w = (float(v) for v in q)
    n = np.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def base_matrix(position=(0, 0, 0), orientation=(0, 0, 0, 1)):
    T = np.eye(4)
    T[:3, :3] = quat_to_matrix(orientation)
    T[:3, 3] = position
    return T


def mdh(a, d, alpha, theta):
    """modified_homogeneous_transform, robots/panda.py:111-122 (theta may be an array)."""
    theta = np.asarray(theta, dtype=np.float64)
    ct, st = np.cos(theta), np.sin(theta)
    ca, sa = np.cos(alpha), np.sin(alpha)
    T = np.zeros(theta.shape + (4, 4))
    T[..., 0, 0] = ct
    T[..., 0, 1] = -st
    T[..., 0, 3] = a
    T[..., 1, 0] = st * ca
    T[..., 1, 1] = ct * ca
    T[..., 1, 2] = -sa
    T[..., 1, 3] = -sa * d
    T[..., 2, 0] = st * sa
    T[..., 2, 1] = ct * sa
    T[..., 2, 2] = ca
    T[..., 2, 3] = ca * d
    T[..., 3, 3] = 1.0
    return T


def panda_frames(q, base=None):
    """All 8 DH frame poses [N, 8, 4, 4] (forward_kinematics loop, panda.py:124-144)."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    T = np.broadcast_to(np.eye(4) if base is None else base, (q.shape[0], 4, 4)).copy()
    out = np.empty((q.shape[0], 8, 4, 4))
    for i, (a, d, alpha) in enumerate(PANDA_DH):
        Ti = mdh(a, d, alpha, q[:, i] if i < 7 else np.zeros(q.shape[0]))
        T = np.matmul(T, Ti)
        out[:, i] = T
    return out


def panda_positions(q, base=None):
    """positions_from_fk, panda.py:151-173 -> [N, 8, 3]."""
    return panda_frames(q, base)[:, :, :3, 3]


def panda_pose_distance(q1, q2, base=None):
    """task_space_pose_distance, panda.py:248-254: sum over the 8 frames, in order."""
    p1, p2 = panda_positions(q1, base), panda_positions(q2, base)
    d = np.zeros(p1.shape[0])
    for i in range(8):
        d = d + np.sqrt(((p1[:, i] - p2[:, i]) ** 2).sum(-1))
    return d


def rpy_matrix(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def origin_matrix(xyz, rpy):
    T = np.eye(4)
    T[:3, :3] = rpy_matrix(rpy)
    T[:3, 3] = xyz
    return T


def chain_frames(origins, q, base=None):
    """Generic URDF chain: frame f = base * prod_{j<f} origin[j] * Rz(q_j); the last
    entry of `origins` (index n) is the fixed tip.  -> [N, n+1, 4, 4]."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    n = q.shape[1]
    T = np.broadcast_to(np.eye(4) if base is None else base, (q.shape[0], 4, 4)).copy()
    out = np.empty((q.shape[0], n + 1, 4, 4))
    for j in range(n + 1):
        O = np.asarray(origins[j], dtype=np.float64)
        if j < n:
            c, s = np.cos(q[:, j]), np.sin(q[:, j])
            Rz = np.zeros((q.shape[0], 4, 4))
            Rz[:, 0, 0] = c
            Rz[:, 0, 1] = -s
            Rz[:, 1, 0] = s
            Rz[:, 1, 1] = c
            Rz[:, 2, 2] = 1
            Rz[:, 3, 3] = 1
            T = np.matmul(np.matmul(T, O), Rz)
        else:
            T = np.matmul(T, O)
        out[:, j] = T
    return out
