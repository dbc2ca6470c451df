"""Thin Python layer over the C ABI: worlds (device tables) and batched kernels.

torch is used only as plumbing (device buffers, the current CUDA stream); every
computation is a call into libmmmp_b200.so.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import (CHECK_BOUNDS, CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF, METRIC_GROUP_L2_SUM, METRIC_L2,
                   METRIC_SQ_SUM, Metric, MmmpError)

GEOMETRY_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "geometry")


# ----------------------------------------------------------------- geometry
def load_model(name: str) -> dict:
    with open(os.path.join(GEOMETRY_DIR, name + ".json")) as f:
        return json.load(f)


def quat_to_matrix(q) -> np.ndarray:
    """xyzw quaternion -> 3x3 (what scipy Rotation.from_quat(q).as_matrix() returns,
    reference robots/panda.py:127)."""
    x, y, z, w = (float(v) for v in q)
    n = np.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def base_matrix(position=(0, 0, 0), orientation=(0, 0, 0, 1)) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = quat_to_matrix(orientation)
    T[:3, 3] = np.asarray(position, dtype=np.float64)
    return T


def mdh_origin(a: float, d: float, alpha: float) -> np.ndarray:
    """Fixed part of the modified-DH transform (theta = 0) of robots/panda.py:111-122:
    T_i(theta) = origin * Rz(theta)."""
    ca, sa = np.cos(alpha), np.sin(alpha)
    return np.array([[1.0, 0.0, 0.0, a], [0.0, ca, -sa, -sa * d], [0.0, sa, ca, ca * d], [0, 0, 0, 1.0]])


def rpy_origin(xyz, rpy) -> np.ndarray:
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = xyz
    return T


def model_origins(model: dict) -> list:
    if "dh_modified" in model:
        return [mdh_origin(r["a"], r["d"], r["alpha"]) for r in model["dh_modified"]]
    org = [rpy_origin(j["xyz"], j["rpy"]) for j in model["joint_origins"]]
    org.append(rpy_origin(model["tip"]["xyz"], model["tip"]["rpy"]))
    return org


@dataclass
class RobotSpec:
    """One serial arm of a spatial world."""
    model: dict
    base: np.ndarray = field(default_factory=lambda: np.eye(4))
    dof_offset: int = 0
    use_pair_tables: bool = True

    def desc(self) -> _lib.RobotDesc:
        m = self.model
        d = _lib.RobotDesc()
        n = int(m["n_joints"])
        d.n_joints = n
        d.dof_offset = int(self.dof_offset)
        for i, v in enumerate(np.asarray(self.base, dtype=np.float64)[:3, :].ravel()):
            d.base[i] = v
        for j, O in enumerate(model_origins(m)):
            for i, v in enumerate(np.asarray(O, dtype=np.float64)[:3, :].ravel()):
                d.origin[j][i] = v
        links = m["links"]
        d.n_links = len(links)
        if len(links) > _lib.MAX_LINKS:
            raise MmmpError("too many collision links")
        sph = sorted(m["spheres"], key=lambda s: s["link_id"])
        if len(sph) > _lib.MAX_SPHERES:
            raise MmmpError(f"{len(sph)} spheres > {_lib.MAX_SPHERES}")
        d.n_spheres = len(sph)
        frames = {}
        for k, s in enumerate(sph):
            for i in range(3):
                d.sphere[k][i] = s["center"][i]
            d.sphere[k][3] = s["radius"]
            d.sphere_link[k] = s["link_id"]
            frames[s["link_id"]] = s["frame"]
        for l in range(len(links)):
            d.link_frame[l] = frames.get(l, 0)
        # link pairs PROVEN collision free over the whole joint range (tools/prove_pairs.py) are
        # not tested at all; with use_pair_tables=False the plain sphere model is used throughout
        proven = {tuple(p) for p in m.get("proven_free_pairs", [])} if self.use_pair_tables else set()
        for a, b in m["self_pairs"]:
            if (a, b) not in proven and (b, a) not in proven:
                d.self_mask[a] |= 1 << b
        tables = m.get("pair_tables", []) if self.use_pair_tables else []
        d.n_pair_tables = len(tables)
        for t, pt in enumerate(tables):
            T = d.pair_table[t]
            for key in ("link_a", "link_b", "joint_u", "joint_v", "n_u", "n_v"):
                setattr(T, key, int(pt[key]))
            for key in ("lo_u", "hi_u", "lo_v", "hi_v"):
                setattr(T, key, float(pt[key]))
            words = np.frombuffer(bytes.fromhex(pt["words_hex"]), dtype=np.uint32)
            C.memmove(T.bits, words.ctypes.data, words.nbytes)
        return d


# obstacle primitives (plain dicts so the oracle can read the same description)
def plane(normal=(0, 0, 1), offset=0.0, body=None) -> dict:
    n = np.asarray(normal, dtype=np.float64)
    n = n / np.linalg.norm(n)
    return {"type": "plane", "normal": n.tolist(), "offset": float(offset), "body": body}


def box(center, half, R=None, body=None) -> dict:
    R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
    return {"type": "box", "center": [float(v) for v in center], "half": [float(v) for v in half],
            "R": R.tolist(), "body": body}


def sphere(center, radius, body=None) -> dict:
    return {"type": "sphere", "center": [float(v) for v in center], "radius": float(radius), "body": body}


def cylinder(center, axis, half_len, radius, body=None) -> dict:
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    return {"type": "cylinder", "center": [float(v) for v in center], "axis": a.tolist(),
            "half_len": float(half_len), "radius": float(radius), "body": body}


def _obstacle_desc(ob: dict, index: int) -> _lib.ObstacleDesc:
    d = _lib.ObstacleDesc()
    d.body = index if ob.get("body") is None else int(ob["body"])
    t = ob["type"]
    if t == "plane":
        d.type = _lib.OBS_PLANE
        vals = list(ob["normal"]) + [ob["offset"]]
    elif t == "sphere":
        d.type = _lib.OBS_SPHERE
        vals = list(ob["center"]) + [ob["radius"]]
    elif t == "box":
        d.type = _lib.OBS_BOX
        vals = list(ob["center"]) + list(ob["half"]) + list(np.asarray(ob["R"], dtype=np.float64).ravel())
    elif t == "cylinder":
        d.type = _lib.OBS_CYLINDER
        vals = list(ob["center"]) + list(ob["axis"]) + [ob["half_len"], ob["radius"]]
    else:
        raise MmmpError(f"unknown obstacle type {t!r}")
    for i, v in enumerate(vals):
        d.p[i] = float(v)
    return d


# ----------------------------------------------------------------- helpers
def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_EMPTY = {}   # device -> a 16-byte buffer whose address stands in for empty tensors


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    """Device address of a tensor; NULL only for None.  torch gives empty tensors a NULL
    data_ptr, which the C ABI would take for a missing argument, so they get the address of a
    small live buffer instead (never dereferenced: their row count is 0)."""
    if t is None:
        return C.c_void_p(0)
    if t.numel() == 0 and t.is_cuda:
        if t.device not in _EMPTY:
            _EMPTY[t.device] = torch.zeros(16, dtype=torch.uint8, device=t.device)
        return C.c_void_p(_EMPTY[t.device].data_ptr())
    return C.c_void_p(t.data_ptr())


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise MmmpError("no CUDA device: mmmp_b200 has no CPU fallback for the roadmap-construction path")
    return torch.device("cuda", torch.cuda.current_device())


def _dev(t: torch.Tensor, dtype) -> torch.Tensor:
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise MmmpError(f"expected a contiguous CUDA tensor of {dtype}, got {t.dtype} on {t.device}")
    return t


def pad4(d: int) -> int:
    """Row width of fp32 configuration / feature rows: 4, 8, 16 or 32 floats (the widths the
    neighbour-search kernels are instantiated for)."""
    for w in (4, 8, 16, 32):
        if d <= w:
            return w
    raise MmmpError(f"row of {d} floats is wider than 32")


def to_device_f64(a) -> torch.Tensor:
    dev = _require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)


def pack_f32(q64: torch.Tensor, ld32: Optional[int] = None) -> torch.Tensor:
    q64 = _dev(q64, torch.float64)
    N, D = q64.shape
    ld32 = pad4(D) if ld32 is None else ld32
    out = torch.empty((N, ld32), dtype=torch.float32, device=q64.device)
    _lib.call("mmmp_pack_f32", _ptr(q64), D, D, N, _ptr(out), ld32, _stream())
    return out


def sample_uniform(lo, hi, n: int, seed: int, first_index: int = 0, round_dp: int = 2):
    """-> (q64 [n, D], q32 [n, pad4(D)]) device tensors."""
    dev = _require_cuda()
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    D = lo.shape[0]
    q64 = torch.empty((n, D), dtype=torch.float64, device=dev)
    q32 = torch.empty((n, pad4(D)), dtype=torch.float32, device=dev)
    _lib.call("mmmp_sample_uniform", lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p), D, n,
              C.c_uint64(seed & (2 ** 64 - 1)), C.c_uint64(first_index), round_dp, _ptr(q64), D, _ptr(q32),
              pad4(D), _stream())
    return q64, q32


def compact_free(flags: torch.Tensor) -> torch.Tensor:
    """Ascending indices of the rows whose flag is 0 (collision free)."""
    flags = _dev(flags, torch.uint8)
    n = flags.numel()
    idx = torch.empty((n,), dtype=torch.int32, device=flags.device)
    cnt = torch.zeros((1,), dtype=torch.int32, device=flags.device)
    _lib.call("mmmp_compact_free", _ptr(flags), n, _ptr(idx), _ptr(cnt), _stream())
    return idx[: int(cnt.item())]


def gather_rows(t: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    idx = _dev(idx, torch.int32)
    if not t.is_contiguous():
        raise MmmpError("gather_rows needs a contiguous tensor")
    row_bytes = t.shape[1] * t.element_size()
    out = torch.empty((idx.numel(), t.shape[1]), dtype=t.dtype, device=t.device)
    _lib.call("mmmp_gather_rows", _ptr(t), row_bytes, _ptr(idx), idx.numel(), _ptr(out), _stream())
    return out


def exclusive_scan(counts: torch.Tensor) -> torch.Tensor:
    counts = _dev(counts, torch.int32)
    out = torch.empty((counts.numel() + 1,), dtype=torch.int64, device=counts.device)
    _lib.call("mmmp_exclusive_scan_i32", _ptr(counts), counts.numel(), _ptr(out), _stream())
    return out


def make_metric(kind: int, dim: int, weights: Sequence[float] = ()) -> Metric:
    m = Metric()
    m.kind, m.dim, m.n_groups = kind, dim, len(weights)
    for i, w in enumerate(weights):
        m.weight[i] = w
    return m


@dataclass
class Features:
    """Row sets of one point set for neighbour search.
    filt: fp32 filter rows [n, 4|8|16|32] whose L2 distance lower-bounds the metric;
    x32 : fp32 exact rows the fp32 metric (m32) is evaluated on;
    x64 : fp64 rows the float64 re-rank metric (m64) is evaluated on."""
    filt: torch.Tensor
    x32: torch.Tensor
    x64: torch.Tensor
    m32: Metric
    m64: Metric
    fdim: int = 0            # filter columns used (0 = m32.dim)

    def rows(self, lo: int, hi: int) -> "Features":
        return Features(self.filt[lo:hi], self.x32[lo:hi], self.x64[lo:hi], self.m32, self.m64, self.fdim)

    def take(self, index: torch.Tensor) -> "Features":
        """The rows `index` (int tensor), gathered into contiguous row sets."""
        i = index.long()
        same = self.filt.data_ptr() == self.x32.data_ptr()
        x32 = self.x32[i].contiguous()
        return Features(x32 if same else self.filt[i].contiguous(), x32, self.x64[i].contiguous(), self.m32, self.m64,
                        self.fdim)

    def __len__(self):
        return self.x32.shape[0]


def l2_features(q64: torch.Tensor, q32: Optional[torch.Tensor] = None, squared: bool = False) -> Features:
    """np.linalg.norm(q1 - q2) on configuration rows (decoupled_prm.py:824, coupled_prm.py:131);
    squared=True gives sum of squares on the rows (planar_arm.py:54-61 on stacked link positions,
    then m64 groups the columns in pairs)."""
    q64 = _dev(q64, torch.float64)
    q32 = pack_f32(q64) if q32 is None else _dev(q32, torch.float32)
    D = q64.shape[1]
    if squared:
        m32 = make_metric(METRIC_SQ_SUM, D)
        m64 = make_metric(METRIC_SQ_SUM, D, [1.0] * (D // 2))
    else:
        m32 = m64 = make_metric(METRIC_L2, D)
    return Features(q32, q32, q64, m32, m64)


def knn(ref: Features, query: Features, k: int, exclude_self: bool = True, causal: bool = False,
        query_id0: int = 0):
    """fp32 scan -> (idx [nq, k] int32, dist [nq, k] fp32), ascending (dist, id)."""
    fr, fq = _dev(ref.filt, torch.float32), _dev(query.filt, torch.float32)
    xr, xq = _dev(ref.x32, torch.float32), _dev(query.x32, torch.float32)
    nq = fq.shape[0]
    idx = torch.empty((nq, k), dtype=torch.int32, device=fr.device)
    dist = torch.empty((nq, k), dtype=torch.float32, device=fr.device)
    fdim = ref.fdim or ref.m32.dim
    _lib.call("mmmp_knn", _ptr(fr), _ptr(xr), fr.shape[0], _ptr(fq), _ptr(xq), nq, fr.shape[1], fdim, xr.shape[1],
              C.byref(ref.m32), k, int(exclude_self), int(causal), query_id0, _ptr(idx), _ptr(dist), _stream())
    return idx, dist


def knn_shard(ref: Features, k: int, shard: int, n_shards: int, exclude_self: bool = True):
    """The slice `shard` of `n_shards` of a search of `ref` against itself, cut in the library's
    spatial scan order (mmmp_knn_shard) -> (rows [m] int32: original row of every answered query,
    idx [m, k] int32, dist [m, k] fp32).  The shards together answer every row exactly once."""
    fr, xr = _dev(ref.filt, torch.float32), _dev(ref.x32, torch.float32)
    n = fr.shape[0]
    cap = (-(-(-(-n // 128)) // n_shards) + 1) * 128
    rows = torch.empty((cap,), dtype=torch.int32, device=fr.device)
    idx = torch.empty((cap, k), dtype=torch.int32, device=fr.device)
    dist = torch.empty((cap, k), dtype=torch.float32, device=fr.device)
    m = C.c_int64(0)
    fdim = ref.fdim or ref.m32.dim
    _lib.call("mmmp_knn_shard", _ptr(fr), _ptr(xr), n, fr.shape[1], fdim, xr.shape[1], C.byref(ref.m32), k,
              int(exclude_self), shard, n_shards, _ptr(rows), C.byref(m), _ptr(idx), _ptr(dist), _stream())
    return rows[:m.value], idx[:m.value], dist[:m.value]


def knn_range(ref: Features, k: int, frac_lo: float, frac_hi: float, exclude_self: bool = True):
    """knn_shard with explicit cuts: the query blocks [floor(B frac_lo), floor(B frac_hi)) of the
    B = ceil(n / 128) blocks of the scan order (mmmp_knn_range) -> (rows, idx, dist)."""
    fr, xr = _dev(ref.filt, torch.float32), _dev(ref.x32, torch.float32)
    n = fr.shape[0]
    blocks = -(-n // 128)
    cap = (int(np.ceil(blocks * max(frac_hi - frac_lo, 0.0))) + 2) * 128
    rows = torch.empty((cap,), dtype=torch.int32, device=fr.device)
    idx = torch.empty((cap, k), dtype=torch.int32, device=fr.device)
    dist = torch.empty((cap, k), dtype=torch.float32, device=fr.device)
    m = C.c_int64(0)
    fdim = ref.fdim or ref.m32.dim
    _lib.call("mmmp_knn_range", _ptr(fr), _ptr(xr), n, fr.shape[1], fdim, xr.shape[1], C.byref(ref.m32), k,
              int(exclude_self), float(frac_lo), float(frac_hi), _ptr(rows), C.byref(m), _ptr(idx), _ptr(dist),
              _stream())
    return rows[:m.value], idx[:m.value], dist[:m.value]


def knn_refine(ref64: torch.Tensor, query64: torch.Tensor, metric64: Metric, cand_idx: torch.Tensor,
               cand_dist: torch.Tensor, k: int, tie_eps: float = 1e-6):
    """fp64 re-rank -> (idx [nq, k], dist64 [nq, k], ambiguous [nq] u8)."""
    ref64, query64 = _dev(ref64, torch.float64), _dev(query64, torch.float64)
    nq, kc = cand_idx.shape
    idx = torch.empty((nq, k), dtype=torch.int32, device=ref64.device)
    dist = torch.empty((nq, k), dtype=torch.float64, device=ref64.device)
    amb = torch.empty((nq,), dtype=torch.uint8, device=ref64.device)
    _lib.call("mmmp_knn_refine", _ptr(ref64), _ptr(query64), ref64.shape[1], C.byref(metric64),
              _ptr(_dev(cand_idx, torch.int32)), _ptr(_dev(cand_dist, torch.float32)), nq, kc, k,
              C.c_double(tie_eps), _ptr(idx), _ptr(dist), _ptr(amb), _stream())
    return idx, dist, amb


def pack_neighbour_records(rows: torch.Tensor, idx: torch.Tensor, dist: torch.Tensor, amb: Optional[torch.Tensor],
                           cap: int) -> torch.Tensor:
    """Rows answered by a sharded search as records [cap, 2 + 3k] int32 (row id, ambiguous, idx,
    dist as fp64 bit pairs; padding records carry row id -1): one all-gather ships them."""
    m, k = idx.shape
    rec = torch.empty((cap, 2 + 3 * k), dtype=torch.int32, device=idx.device)
    _lib.call("mmmp_pack_neighbour_records", _ptr(_dev(rows, torch.int32)), _ptr(_dev(idx, torch.int32)),
              _ptr(_dev(dist, torch.float64)), _ptr(amb), m, k, cap, _ptr(rec), _stream())
    return rec


def scatter_neighbour_records(rec: torch.Tensor, k: int, n: int, want_dist: bool = True):
    """Gathered records -> (idx [n, k] int32, dist [n, k] fp64 or None, amb [n] u8) in row order."""
    rec = _dev(rec, torch.int32)
    idx = torch.empty((n, k), dtype=torch.int32, device=rec.device)
    dist = torch.empty((n, k), dtype=torch.float64, device=rec.device) if want_dist else None
    amb = torch.empty((n,), dtype=torch.uint8, device=rec.device)
    _lib.call("mmmp_scatter_neighbour_records", _ptr(rec), rec.shape[0], k, n, _ptr(idx), _ptr(dist), _ptr(amb),
              _stream())
    return idx, dist, amb


def knn_exact(ref: Features, query: Features, k: int, pad: int = 6, exclude_self: bool = True,
              causal: bool = False, query_id0: int = 0):
    """scan with k+pad candidates, then fp64 re-rank -> (idx, dist64, ambiguous)."""
    kc = min(k + pad, _lib.MAX_K)
    ci, cd = knn(ref, query, kc, exclude_self, causal, query_id0)
    idx, dist, amb = knn_refine(ref.x64, query.x64, ref.m64, ci, cd, k)
    kc_wide = min(k + 16, _lib.MAX_K - 1)
    if not causal and kc < kc_wide and bool(amb.any()):
        # rows whose rank k the fp32 scan could not separate from its last candidate: searched again,
        # alone, with k + 16 candidates (self comes back at distance 0 and is dropped here)
        bad = amb.nonzero().flatten()
        sub = query.take(bad)
        ci2, cd2 = knn(ref, sub, kc_wide + 1, exclude_self=False)
        if exclude_self:
            ids = (bad + query_id0).to(torch.int32)
            keep = ci2 != ids[:, None]
            order = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)[:, :kc_wide]
            ci2, cd2 = torch.gather(ci2, 1, order).contiguous(), torch.gather(cd2, 1, order).contiguous()
        else:
            ci2, cd2 = ci2[:, :kc_wide].contiguous(), cd2[:, :kc_wide].contiguous()
        idx[bad], dist[bad], amb[bad] = knn_refine(ref.x64, sub.x64, ref.m64, ci2, cd2, k)
    return idx, dist, amb


def radius_search(ref: Features, query: Features, r: float, strict: bool = False, exclude_self: bool = True,
                  exclude_zero: bool = False, causal: bool = False, query_id0: int = 0):
    """CSR neighbour lists within r: (offsets [nq+1] int64, idx int32, dist fp64); ids ascending per row."""
    ref32, query32 = _dev(ref.x32, torch.float32), _dev(query.x32, torch.float32)
    ref64, query64 = _dev(ref.x64, torch.float64), _dev(query.x64, torch.float64)
    nq, ldf = query32.shape
    counts = torch.zeros((nq,), dtype=torch.int32, device=ref32.device)
    args = lambda cnt, off, oi, od: (  # noqa: E731
        _ptr(ref32), _ptr(ref64), ref32.shape[0], _ptr(query32), _ptr(query64), nq, ldf, ref64.shape[1],
        C.byref(ref.m32), C.byref(ref.m64), C.c_double(r), int(strict), int(exclude_self), int(exclude_zero),
        int(causal), query_id0, _ptr(cnt), _ptr(off), _ptr(oi), _ptr(od), _stream())
    _lib.call("mmmp_radius", *args(counts, None, None, None))
    offsets = exclusive_scan(counts)
    total = int(offsets[-1].item())
    idx = torch.empty((max(total, 1),), dtype=torch.int32, device=ref32.device)
    dist = torch.empty((max(total, 1),), dtype=torch.float64, device=ref32.device)
    if total > 0:
        _lib.call("mmmp_radius", *args(None, offsets, idx, dist))
    return offsets, idx[:total], dist[:total]


def greedy_degree_connect(nbr_idx: np.ndarray, edge_free: np.ndarray, k2: int, cap: Optional[int] = None):
    """Native host replay of connect_nearest_neighbors_degree (decoupled_prm.py:496-503)."""
    nbr_idx = np.ascontiguousarray(nbr_idx, dtype=np.int32)
    edge_free = np.ascontiguousarray(edge_free, dtype=np.uint8)
    N, k1 = nbr_idx.shape
    cap = k2 if cap is None else cap
    adj = np.empty((N, cap), dtype=np.int32)
    deg = np.empty((N,), dtype=np.int32)
    _lib.call("mmmp_greedy_degree_connect", nbr_idx.ctypes.data_as(C.c_void_p), edge_free.ctypes.data_as(C.c_void_p),
              N, k1, k2, cap, adj.ctypes.data_as(C.c_void_p), deg.ctypes.data_as(C.c_void_p))
    return adj, deg


def greedy_degree_connect_device(nbr_idx: torch.Tensor, edge_free: torch.Tensor, k2: int, cap: Optional[int] = None,
                                 return_rounds: bool = False):
    """connect_nearest_neighbors_degree's accept pass (decoupled_prm.py:496-503) on the GPU:
    device tables in, (adj [N, cap] int32 -1 padded, deg [N] int32) out, identical to the
    sequential replay (mmmp_greedy_degree_connect), list order included."""
    nbr_idx = _dev(nbr_idx, torch.int32)
    edge_free = _dev(edge_free, torch.uint8)
    N, k1 = nbr_idx.shape
    if tuple(edge_free.shape) != (N, k1):
        raise MmmpError("greedy_degree_connect_device: edge_free must have the shape of nbr_idx")
    cap = k2 if cap is None else cap
    adj = torch.empty((N, cap), dtype=torch.int32, device=nbr_idx.device)
    deg = torch.empty((N,), dtype=torch.int32, device=nbr_idx.device)
    rounds = torch.zeros((1,), dtype=torch.int32, device=nbr_idx.device) if return_rounds else None
    _lib.call("mmmp_greedy_degree_connect_device", _ptr(nbr_idx), _ptr(edge_free), N, k1, k2, cap, _ptr(adj), _ptr(deg),
              _ptr(rounds), _stream())
    return (adj, deg, int(rounds.item())) if return_rounds else (adj, deg)


def connected_components(adj: torch.Tensor, mask: Optional[torch.Tensor] = None, sync: bool = True):
    """Partition of compute_connected_components (decoupled_prm.py:613-636) on the GPU.
    adj [N, cap] int32 neighbour table (-1 padded), mask [N, cap] uint8 (1 = edge present) or None.
    -> (labels [N] int32: smallest node id of the node's component, number of components);
    sync=False leaves the count on the device (a [1] int64 tensor) and does not wait for the stream."""
    dev = _require_cuda()
    adj = _dev(adj, torch.int32)
    n, cap = adj.shape
    if mask is not None:
        mask = _dev(mask, torch.uint8)
        if tuple(mask.shape) != (n, cap):
            raise MmmpError("connected_components: mask must have the shape of adj")
    labels = torch.empty((n,), dtype=torch.int32, device=dev)
    count = torch.zeros((1,), dtype=torch.int64, device=dev)
    _lib.call("mmmp_connected_components", _ptr(adj), _ptr(mask) if mask is not None else None, n, cap, _ptr(labels),
              _ptr(count), _stream())
    return labels, (int(count.item()) if sync else count)


def launch_count() -> int:
    return int(_lib.load().mmmp_launch_count())


# ----------------------------------------------------------------- worlds
class _World:
    def __init__(self):
        self._h = C.c_void_p(0)

    @property
    def handle(self) -> C.c_void_p:
        if not self._h:
            raise MmmpError("world already destroyed")
        return self._h

    def close(self):
        if self._h:
            _lib.load().mmmp_world_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # collision ----------------------------------------------------------------
    _qdtype = torch.float32

    def collide_configs(self, q: torch.Tensor, robot: int = -1, flags: int = CHECK_SELF | CHECK_OBSTACLES):
        q = _dev(q, self._qdtype)
        out = torch.empty((q.shape[0],), dtype=torch.uint8, device=q.device)
        _lib.call("mmmp_collide_configs", self.handle, robot, _ptr(q), q.shape[1], q.shape[0], flags, _ptr(out),
                  _stream())
        return out

    def collide_edges(self, q: torch.Tensor, edges: torch.Tensor, n_steps: int, step: float, robot: int = -1,
                      flags: int = CHECK_SELF | CHECK_OBSTACLES):
        q = _dev(q, self._qdtype)
        edges = _dev(edges, torch.int32)
        E = edges.numel() // 2
        out = torch.empty((E,), dtype=torch.uint8, device=q.device)
        _lib.call("mmmp_collide_edges", self.handle, robot, _ptr(q), q.shape[1], _ptr(edges), E, n_steps,
                  C.c_double(step), flags, _ptr(out), _stream())
        return out

    def collide_segments(self, qa: torch.Tensor, qb: torch.Tensor, n_steps: int, step: float, robot: int = -1,
                         flags: int = CHECK_SELF | CHECK_OBSTACLES):
        qa, qb = _dev(qa, self._qdtype), _dev(qb, self._qdtype)
        out = torch.empty((qa.shape[0],), dtype=torch.uint8, device=qa.device)
        _lib.call("mmmp_collide_segments", self.handle, robot, _ptr(qa), _ptr(qb), qa.shape[1], qa.shape[0],
                  n_steps, C.c_double(step), flags, _ptr(out), _stream())
        return out


def edges_from_knn(nbr: torch.Tensor) -> torch.Tensor:
    nbr = _dev(nbr, torch.int32)
    n, k = nbr.shape
    edges = torch.empty((n * k, 2), dtype=torch.int32, device=nbr.device)
    _lib.call("mmmp_edges_from_knn", _ptr(nbr), n, k, _ptr(edges), _stream())
    return edges


class SpatialWorld(_World):
    """Robots (serial chains + sphere decomposition) and obstacle primitives on the device."""

    def __init__(self, robots: Sequence[RobotSpec], obstacles: Sequence[dict] = ()):
        super().__init__()
        _require_cuda()
        self.robots = list(robots)
        self.obstacles = list(obstacles)
        rd = (_lib.RobotDesc * len(self.robots))(*[r.desc() for r in self.robots])
        od = (_lib.ObstacleDesc * max(1, len(self.obstacles)))(*[_obstacle_desc(o, i) for i, o in
                                                                  enumerate(self.obstacles)])
        h = C.c_void_p(0)
        _lib.call("mmmp_world_create", rd, len(self.robots), od, len(self.obstacles), C.byref(h))
        self._h = h

    @property
    def dof(self) -> int:
        return int(_lib.load().mmmp_world_dof(self.handle))

    def collide_knn_edges(self, q32: torch.Tensor, nbr: torch.Tensor, n_steps: int, step: float, robot: int = -1,
                          flags: int = CHECK_SELF | CHECK_OBSTACLES, rows=None, share_across_rows: bool = False):
        """Verdicts of a candidate table: free [hi - lo, k] uint8 for rows [lo, hi) of nbr [n, k]
        (1 = edge row -> neighbour collision free).  A pair of nodes that name each other is evaluated
        once, from the lower id to the higher (mmmp_collide_knn_edges); with share_across_rows the
        slots whose verdict another row block holds come back as 2 -- resolve_knn_edges() on the
        gathered table fills them in."""
        q32, nbr = _dev(q32, torch.float32), _dev(nbr, torch.int32)
        n, k = nbr.shape
        lo, hi = (0, n) if rows is None else rows
        free = torch.empty((hi - lo, k), dtype=torch.uint8, device=nbr.device)
        _lib.call("mmmp_collide_knn_edges", self.handle, robot, _ptr(q32), q32.shape[1], _ptr(nbr), n, k, lo, hi,
                  1 if share_across_rows else 0, n_steps, C.c_double(step), flags, _ptr(free), _stream())
        return free

    @staticmethod
    def resolve_knn_edges(nbr: torch.Tensor, free: torch.Tensor, rows=None):
        """In place: slots of `free` marked 2 take their partner's verdict (rows [lo, hi) of the table)."""
        nbr = _dev(nbr, torch.int32)
        n, k = nbr.shape
        lo, hi = (0, n) if rows is None else rows
        assert free.dtype == torch.uint8 and free.is_contiguous() and free.shape == (hi - lo, k)
        _lib.call("mmmp_resolve_knn_edges", _ptr(nbr), n, k, lo, hi, _ptr(free), _stream())
        return free

    def fk(self, q32: torch.Tensor, robot: int = 0, poses: bool = False):
        q32 = _dev(q32, torch.float32)
        L = self.robots[robot].model["n_joints"] + 1
        pos = torch.empty((q32.shape[0], L, 3), dtype=torch.float32, device=q32.device)
        T = torch.empty((q32.shape[0], L, 12), dtype=torch.float32, device=q32.device) if poses else None
        _lib.call("mmmp_fk_chain", self.handle, robot, _ptr(q32), q32.shape[1], q32.shape[0], _ptr(pos), _ptr(T),
                  _stream())
        return (pos, T) if poses else pos

    def fk64(self, q64: torch.Tensor, robot: int = 0):
        q64 = _dev(q64, torch.float64)
        L = self.robots[robot].model["n_joints"] + 1
        pos = torch.empty((q64.shape[0], L, 3), dtype=torch.float64, device=q64.device)
        _lib.call("mmmp_fk_chain_f64", self.handle, robot, _ptr(q64), q64.shape[1], q64.shape[0], _ptr(pos),
                  _stream())
        return pos

    def ik(self, target_pos: torch.Tensor, target_quat: Optional[torch.Tensor], q0: torch.Tensor, robot: int = 0,
           max_iters: int = 100, damping: float = 0.05, tol_pos: float = 1e-4, tol_rot: float = 1e-3,
           max_step: float = 0.3):
        """Batched damped-least-squares IK of the tip frame (mmmp_ik_dls).  target_pos [N,3] fp64,
        target_quat [N,4] (x,y,z,w) fp64 or None, q0 [N,D] or [1,D] fp64 start configurations.
        -> (q [N,D] fp64, converged [N] uint8, residual [N,2] fp32 = |e_pos|, |e_rot|)."""
        target_pos = _dev(target_pos, torch.float64)
        q0 = _dev(q0, torch.float64)
        N, D = target_pos.shape[0], self.robots[robot].model["n_joints"]
        if target_quat is not None:
            target_quat = _dev(target_quat, torch.float64)
            if tuple(target_quat.shape) != (N, 4):
                raise MmmpError("ik: target_quat must be [N, 4]")
        if target_pos.shape[1] != 3 or q0.shape[1] != D or q0.shape[0] not in (1, N):
            raise MmmpError("ik: target_pos must be [N, 3], q0 [N, D] or [1, D]")
        lim = np.ascontiguousarray(self.robots[robot].model["joint_limits"], dtype=np.float64)
        lo, hi = np.ascontiguousarray(lim[:, 0]), np.ascontiguousarray(lim[:, 1])
        q = torch.empty((N, D), dtype=torch.float64, device=target_pos.device)
        conv = torch.empty((N,), dtype=torch.uint8, device=target_pos.device)
        res = torch.empty((N, 2), dtype=torch.float32, device=target_pos.device)
        _lib.call("mmmp_ik_dls", self.handle, robot, _ptr(target_pos), _ptr(target_quat), _ptr(q0),
                  0 if q0.shape[0] == 1 and N != 1 else D, N, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p),
                  max_iters, damping, tol_pos, tol_rot, max_step, _ptr(q), D, _ptr(conv), _ptr(res), _stream())
        return q, conv, res

    def metric_groups(self, robot: int = 0):
        frames = (C.c_int32 * _lib.MAX_GROUPS)()
        weights = (C.c_float * _lib.MAX_GROUPS)()
        n = C.c_int(0)
        _lib.call("mmmp_fk_metric_groups", self.handle, robot, frames, weights, C.byref(n))
        return list(frames[: n.value]), list(weights[: n.value])

    def fk_features(self, q64: torch.Tensor, q32: Optional[torch.Tensor] = None, robot: int = 0,
                    dual: bool = True) -> Features:
        """Row sets of the FK metric (Panda.task_space_pose_distance, robots/panda.py:248-254):
        exact fp32 rows = xyz of the q-dependent frames (coincident frames merged, weights =
        multiplicity), filter rows = dual rows (centroid + half-chain difference, see
        mmmp_fk_features; dual=False: the centroid alone), fp64 rows = all frame positions."""
        q64 = _dev(q64, torch.float64)
        q32 = pack_f32(q64) if q32 is None else _dev(q32, torch.float32)
        frames, weights = self.metric_groups(robot)
        ldf = pad4(3 * len(frames))
        n = q32.shape[0]
        out = torch.empty((n, ldf), dtype=torch.float32, device=q32.device)
        ldfilt = 8 if dual else 4
        filt = torch.empty((n, ldfilt), dtype=torch.float32, device=q32.device)
        fr = (C.c_int32 * len(frames))(*frames)
        wt = (C.c_float * len(frames))(*weights)
        _lib.call("mmmp_fk_features", self.handle, robot, _ptr(q32), q32.shape[1], n, fr, wt, len(frames),
                  _ptr(out), ldf, _ptr(filt), ldfilt, _stream())
        L = self.robots[robot].model["n_joints"] + 1
        pos64 = self.fk64(q64, robot).reshape(n, 3 * L)
        m32 = make_metric(METRIC_GROUP_L2_SUM, 3 * len(frames), weights)
        m64 = make_metric(METRIC_GROUP_L2_SUM, 3 * L, [1.0] * L)
        return Features(filt, out, pos64, m32, m64, 6 if dual else 3)

    def collide_robot_pairs(self, qa: torch.Tensor, qb: torch.Tensor, ra: int, rb: int):
        qa, qb = _dev(qa, torch.float32), _dev(qb, torch.float32)
        out = torch.empty((qa.shape[0],), dtype=torch.uint8, device=qa.device)
        _lib.call("mmmp_collide_robot_pairs", self.handle, ra, rb, _ptr(qa), _ptr(qb), qa.shape[1], qa.shape[0],
                  _ptr(out), _stream())
        return out

    def roadmap_candidates_host(self, q64: np.ndarray, robot: int, metric_kind: int, k: int, n_steps: int,
                                step: float, flags: int = CHECK_SELF | CHECK_OBSTACLES, pinned: bool = False):
        """Host arrays in, host arrays out (the e2e call): see include/mmmp_b200.h.
        pinned=True returns the outputs in page-locked host memory (views of torch tensors)."""
        q64 = np.ascontiguousarray(q64, dtype=np.float64)
        N, D = q64.shape

        def host(shape, np_dtype, t_dtype):
            if pinned:
                return torch.empty(shape, dtype=t_dtype, pin_memory=True).numpy()
            return np.empty(shape, dtype=np_dtype)

        node_src = host((N,), np.int32, torch.int32)
        nbr = host((N, k), np.int32, torch.int32)
        dist = host((N, k), np.float64, torch.float64)
        free = host((N, k), np.uint8, torch.uint8)
        n_nodes = C.c_int64(0)
        timings = np.zeros((8,), dtype=np.float32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        _lib.call("mmmp_roadmap_candidates_host", self.handle, robot, vp(q64), N, D, metric_kind, k, n_steps,
                  C.c_double(step), flags, C.byref(n_nodes), vp(node_src), vp(nbr), vp(dist), vp(free), vp(timings))
        n = n_nodes.value
        return {"n_nodes": n, "node_src": node_src[:n], "nbr_idx": nbr[:n], "nbr_dist": dist[:n],
                "edge_free": free[:n], "timings_ms": timings}


    def roadmap_build_host(self, q64: np.ndarray, robot: int, metric_kind: int, k1: int, k2: int, n_steps: int,
                           step: float, flags: int = CHECK_SELF | CHECK_OBSTACLES, out: Optional[dict] = None):
        """Host arrays in, the ROADMAP out (mmmp_roadmap_build_host): candidate tables as
        roadmap_candidates_host plus adj [n, k2], deg [n], labels [n], n_components.
        `out` = host_buffers(N, k1, k2) re-uses page-locked output buffers across calls."""
        q64 = np.ascontiguousarray(q64, dtype=np.float64)
        N, D = q64.shape
        b = host_buffers(N, k1, k2) if out is None else out
        if b["nbr_idx"].shape[0] < N or b["nbr_idx"].shape[1] != k1 or b["adj"].shape[1] != k2:
            raise MmmpError("roadmap_build_host: output buffers do not fit this call")
        n_nodes, n_comp = C.c_int64(0), C.c_int64(0)
        timings = np.zeros((10,), dtype=np.float32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        _lib.call("mmmp_roadmap_build_host", self.handle, robot, vp(q64), N, D, metric_kind, k1, k2, n_steps,
                  C.c_double(step), flags, C.byref(n_nodes), vp(b["node_src"]), vp(b["nbr_idx"]), vp(b["nbr_dist"]),
                  vp(b["edge_free"]), vp(b["adj"]), vp(b["deg"]), vp(b["labels"]), C.byref(n_comp), vp(timings))
        n = n_nodes.value
        res = {key: b[key][:n] for key in ("node_src", "nbr_idx", "nbr_dist", "edge_free", "adj", "deg", "labels")}
        res.update(n_nodes=n, n_components=n_comp.value, timings_ms=timings)
        return res


def host_buffers(N: int, k1: int, k2: int, pinned: bool = True) -> dict:
    """Output buffers of SpatialWorld.roadmap_build_host for up to N samples (page-locked by default:
    the library's device-to-host copies then overlap its kernels)."""
    def host(shape, t_dtype):
        return torch.empty(shape, dtype=t_dtype, pin_memory=pinned and torch.cuda.is_available()).numpy()
    return {"node_src": host((N,), torch.int32), "nbr_idx": host((N, k1), torch.int32),
            "nbr_dist": host((N, k1), torch.float64), "edge_free": host((N, k1), torch.uint8),
            "adj": host((N, k2), torch.int32), "deg": host((N,), torch.int32), "labels": host((N,), torch.int32)}


class PlanarWorld(_World):
    """Planar arms, map bounds and Rectangle/Circle obstacles on the device (fp64)."""
    _qdtype = torch.float64

    def __init__(self, arms: Sequence[dict], bounds, obstacles: Sequence[dict] = (), joint_link_radius: float = 0.5):
        """arms: [{links: [..], base: (x, y)}]; obstacles: {type: 'rect', x0,y0,x1,y1} | {type:'circle', cx,cy,r}."""
        super().__init__()
        _require_cuda()
        d = _lib.PlanarDesc()
        d.n_arms = len(arms)
        off = 0
        for a, arm in enumerate(arms):
            links = [float(v) for v in arm["links"]]
            d.n_links[a] = len(links)
            d.dof_offset[a] = off
            off += len(links)
            for l, v in enumerate(links):
                d.link_len[a][l] = v
            d.base[a][0], d.base[a][1] = float(arm["base"][0]), float(arm["base"][1])
        for i, v in enumerate(bounds):
            d.bounds[i] = float(v)
        d.n_obstacles = len(obstacles)
        for o, ob in enumerate(obstacles):
            if ob["type"] == "rect":
                d.obs_type[o] = _lib.POBS_RECT
                vals = [ob["x0"], ob["y0"], ob["x1"], ob["y1"]]
            elif ob["type"] == "circle":
                d.obs_type[o] = _lib.POBS_CIRCLE
                vals = [ob["cx"], ob["cy"], ob["r"], 0.0]
            else:
                raise MmmpError(f"unknown planar obstacle {ob['type']!r}")
            for i, v in enumerate(vals):
                d.obs[o][i] = float(v)
        d.joint_link_radius = float(joint_link_radius)
        self.arms = list(arms)
        h = C.c_void_p(0)
        _lib.call("mmmp_planar_world_create", C.byref(d), C.byref(h))
        self._h = h

    def fk(self, q64: torch.Tensor, arm: int = 0):
        q64 = _dev(q64, torch.float64)
        n = len(self.arms[arm]["links"])
        out = torch.empty((q64.shape[0], n, 2), dtype=torch.float64, device=q64.device)
        _lib.call("mmmp_planar_fk", self.handle, arm, _ptr(q64), q64.shape[1], q64.shape[0], _ptr(out), _stream())
        return out
