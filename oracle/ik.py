"""Damped-least-squares inverse kinematics in float64 numpy (TEST INFRASTRUCTURE).

Restates, step for step, the iteration of mmmp_b200/csrc/ik.cu (mmmp_ik_dls), which stands
in for pybullet.calculateInverseKinematics behind Panda.solve_closest_arm_ik
(robots/panda.py:207-208, utils/ik_utils.py:25-27) in sample_valid_in_task_space
(planners/prm/pybullet/decoupled/decoupled_prm.py:323-344).

PARITY UNPINNED against pybullet: Bullet's solver (pybullet==3.2.6, requirements.txt) is
absent from the reference tree and not installable here, and no reference test records an
IK answer.  What is pinned: the published method (Buss & Kim, selectively damped / damped
least squares on the geometric Jacobian), and the task-space POSITION sampler
sample_task_space_position (:346-369), restated in `task_space_positions` and checked
against the reference's own function (tests/golden/task_space_sampler.npz).
"""
import numpy as np

from . import chain
from .spheres import model_origins


def quat_to_rot(q):
    q = np.asarray(q, dtype=np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    s = 2.0 / (x * x + y * y + z * z + w * w)
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - s * (y * y + z * z); R[..., 0, 1] = s * (x * y - z * w); R[..., 0, 2] = s * (x * z + y * w)
    R[..., 1, 0] = s * (x * y + z * w); R[..., 1, 1] = 1 - s * (x * x + z * z); R[..., 1, 2] = s * (y * z - x * w)
    R[..., 2, 0] = s * (x * z - y * w); R[..., 2, 1] = s * (y * z + x * w); R[..., 2, 2] = 1 - s * (x * x + y * y)
    return R


def task_error(model, base, q, target_pos, target_quat=None):
    """-> (e [N,6], frames [N,n+1,4,4]); e = [p* - p, 0.5 sum_c (R[:,c] x R*[:,c])]."""
    F = chain.chain_frames(model_origins(model), q, base)
    T = F[:, -1]
    e = np.zeros((q.shape[0], 6))
    e[:, :3] = target_pos - T[:, :3, 3]
    if target_quat is not None:
        Rt = quat_to_rot(target_quat)
        for c in range(3):
            e[:, 3:] += 0.5 * np.cross(T[:, :3, c], Rt[:, :, c])
    return e, F


def ik_dls(model, base, target_pos, target_quat, q0, max_iters=100, damping=0.05, tol_pos=1e-4, tol_rot=1e-3,
           max_step=0.3):
    """-> (q [N,n], converged [N] bool, residual [N,2])."""
    target_pos = np.atleast_2d(np.asarray(target_pos, dtype=np.float64))
    N = target_pos.shape[0]
    n = model["n_joints"]
    lim = np.asarray(model["joint_limits"], dtype=np.float64)
    q = np.broadcast_to(np.atleast_2d(np.asarray(q0, dtype=np.float64)), (N, n)).copy()
    if target_quat is not None:
        target_quat = np.broadcast_to(np.atleast_2d(np.asarray(target_quat, dtype=np.float64)), (N, 4))
    done = np.zeros(N, bool)
    res = np.zeros((N, 2))
    for it in range(max_iters + 1):
        live = np.nonzero(~done)[0]
        if len(live) == 0:
            break
        e, F = task_error(model, base, q[live], target_pos[live], None if target_quat is None else target_quat[live])
        res[live, 0] = np.linalg.norm(e[:, :3], axis=1)
        res[live, 1] = np.linalg.norm(e[:, 3:], axis=1)
        ok = (res[live, 0] < tol_pos) & (res[live, 1] < tol_rot)
        done[live[ok]] = True
        if it == max_iters:
            break
        live, e, F = live[~ok], e[~ok], F[~ok]
        if len(live) == 0:
            break
        p_tool = F[:, -1, :3, 3]
        J = np.zeros((len(live), 6, n))
        for j in range(n):
            axis, pj = F[:, j, :3, 2], F[:, j, :3, 3]
            J[:, :3, j] = np.cross(axis, p_tool - pj)
            if target_quat is not None:
                J[:, 3:, j] = axis
        A = np.matmul(J, J.transpose(0, 2, 1)) + (damping ** 2) * np.eye(6)
        if target_quat is None and damping == 0.0:
            A[:, 3:, 3:] = np.eye(3)
        y = np.linalg.solve(A, e[:, :, None])[:, :, 0]
        dq = np.einsum("nrj,nr->nj", J, y)
        big = np.abs(dq).max(1)
        scale = np.where(big > max_step, max_step / np.maximum(big, 1e-300), 1.0)
        q[live] = np.clip(q[live] + scale[:, None] * dq, lim[:, 0], lim[:, 1])
    return q, done, res


def task_space_positions(start_pos, goal_pos, t, chi2, theta, k=1, mu=0.4, sigma=0.2):
    """sample_task_space_position (:346-369) for given draws: t ~ U(0,1), chi2 ~ chisquare(k),
    theta ~ U(0, 2 pi), consumed in that order per sample by the reference."""
    start_pos, goal_pos = np.asarray(start_pos, dtype=np.float64), np.asarray(goal_pos, dtype=np.float64)
    W = goal_pos - start_pos
    W = W / np.linalg.norm(W)
    U = np.array([0.0, 1.0, 0.0]) if np.allclose(W, [1.0, 0.0, 0.0]) else np.array([1.0, 0.0, 0.0])
    U = U - np.dot(U, W) * W
    U = U / np.linalg.norm(U)
    V = np.cross(W, U)
    t, chi2, theta = (np.asarray(a, dtype=np.float64) for a in (t, chi2, theta))
    d = mu / k * (sigma ** 2 / (2 * k)) ** 0.5 * chi2
    return (start_pos + t[:, None] * (goal_pos - start_pos)
            + d[:, None] * (np.cos(theta)[:, None] * U + np.sin(theta)[:, None] * V))
