#!/usr/bin/env python
"""Export the convex-hull vertices of the Panda collision meshes for the CPU hull
oracle (build container only; reads /root/reference).  Output: oracle/data/panda_hulls.npz
(link name -> [n,3] float64 vertices in the frame the sphere model uses, i.e. hand
and fingers already expressed in the panda_link7 frame)."""
import os
import sys

import numpy as np
from scipy.spatial import ConvexHull

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fit_spheres import REF, load_stl, rpy_matrix  # noqa: E402

mesh_dir = os.path.join(REF, "models/franka_panda/meshes/collision")
hand_R = rpy_matrix((0, 0, -0.785398163397))
T_hand = np.eye(4); T_hand[:3, :3] = hand_R; T_hand[2, 3] = 0.107
T_lf = T_hand @ np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0.0584], [0, 0, 0, 1.0]])
T_rf = T_lf @ np.diag([-1.0, -1.0, 1.0, 1.0])
parts = [("panda_link0", "link0", np.eye(4))] + [(f"panda_link{i}", f"link{i}", np.eye(4)) for i in range(1, 8)]
parts += [("panda_hand", "hand", T_hand), ("panda_leftfinger", "finger", T_lf), ("panda_rightfinger", "finger", T_rf)]
out = {}
for name, mesh, T in parts:
    v = load_stl(os.path.join(mesh_dir, mesh + ".stl"))
    v = v[ConvexHull(v).vertices]
    out[name] = v @ T[:3, :3].T + T[:3, 3]
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "data", "panda_hulls.npz")
np.savez_compressed(dst, **out)
print(dst, {k: v.shape for k, v in out.items()})
