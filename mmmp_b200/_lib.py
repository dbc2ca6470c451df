"""ctypes binding of the C-ABI library (include/mmmp_b200.h).

The product path has NO CPU fallback: if the library is missing, or a call
returns a non-zero status (for instance because there is no CUDA device), a
`MmmpError` is raised.  Nothing under `oracle/` is ever imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_JOINTS = 8
MAX_SPHERES = 320
MAX_PAIR_TABLES = 2
PAIR_TABLE_WORDS = 4096
MAX_LINKS = 16
MAX_ROBOTS = 16
MAX_OBSTACLES = 64
MAX_K = 64
MAX_GROUPS = 8

OBS_PLANE, OBS_SPHERE, OBS_BOX, OBS_CYLINDER = 0, 1, 2, 3
POBS_RECT, POBS_CIRCLE = 0, 1
CHECK_SELF, CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_BOUNDS, CHECK_EXHAUSTIVE, CHECK_NO_CAPSULES = 1, 2, 4, 8, 16, 32
METRIC_L2, METRIC_GROUP_L2_SUM, METRIC_SQ_SUM = 0, 1, 2


class MmmpError(RuntimeError):
    pass


class PairTable(C.Structure):
    _fields_ = [
        ("link_a", C.c_int32), ("link_b", C.c_int32),
        ("joint_u", C.c_int32), ("joint_v", C.c_int32),
        ("n_u", C.c_int32), ("n_v", C.c_int32),
        ("lo_u", C.c_double), ("hi_u", C.c_double), ("lo_v", C.c_double), ("hi_v", C.c_double),
        ("bits", C.c_uint32 * PAIR_TABLE_WORDS),
    ]


class RobotDesc(C.Structure):
    _fields_ = [
        ("n_joints", C.c_int32),
        ("n_spheres", C.c_int32),
        ("n_links", C.c_int32),
        ("dof_offset", C.c_int32),
        ("n_pair_tables", C.c_int32),
        ("reserved", C.c_int32),
        ("base", C.c_double * 12),
        ("origin", (C.c_double * 12) * (MAX_JOINTS + 1)),
        ("sphere", (C.c_double * 4) * MAX_SPHERES),
        ("sphere_link", C.c_int32 * MAX_SPHERES),
        ("link_frame", C.c_int32 * MAX_LINKS),
        ("self_mask", C.c_uint32 * MAX_LINKS),
        ("pair_table", PairTable * MAX_PAIR_TABLES),
    ]


class ObstacleDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("body", C.c_int32), ("p", C.c_double * 16)]


class PlanarDesc(C.Structure):
    _fields_ = [
        ("n_arms", C.c_int32),
        ("n_obstacles", C.c_int32),
        ("n_links", C.c_int32 * MAX_ROBOTS),
        ("dof_offset", C.c_int32 * MAX_ROBOTS),
        ("link_len", (C.c_double * MAX_JOINTS) * MAX_ROBOTS),
        ("base", (C.c_double * 2) * MAX_ROBOTS),
        ("bounds", C.c_double * 4),
        ("obs_type", C.c_int32 * MAX_OBSTACLES),
        ("obs", (C.c_double * 4) * MAX_OBSTACLES),
        ("joint_link_radius", C.c_double),
    ]


class Metric(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("dim", C.c_int32),
        ("n_groups", C.c_int32),
        ("weight", C.c_float * MAX_GROUPS),
    ]


_P = C.c_void_p
_I = C.c_int
_I64 = C.c_int64
_U64 = C.c_uint64
_U32 = C.c_uint32
_D = C.c_double

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "mmmp_world_create": [C.POINTER(RobotDesc), _I, C.POINTER(ObstacleDesc), _I, C.POINTER(_P)],
    "mmmp_planar_world_create": [C.POINTER(PlanarDesc), C.POINTER(_P)],
    "mmmp_world_destroy": [_P],
    "mmmp_world_dof": [_P],
    "mmmp_sample_uniform": [_P, _P, _I, _I64, _U64, _U64, _I, _P, _I, _P, _I, _P],
    "mmmp_pack_f32": [_P, _I, _I, _I64, _P, _I, _P],
    "mmmp_fk_chain": [_P, _I, _P, _I, _I64, _P, _P, _P],
    "mmmp_fk_features": [_P, _I, _P, _I, _I64, _P, _P, _I, _P, _I, _P, _I, _P],
    "mmmp_fk_metric_groups": [_P, _I, _P, _P, _P],
    "mmmp_fk_chain_f64": [_P, _I, _P, _I, _I64, _P, _P],
    "mmmp_planar_fk": [_P, _I, _P, _I, _I64, _P, _P],
    "mmmp_collide_configs": [_P, _I, _P, _I, _I64, _U32, _P, _P],
    "mmmp_collide_edges": [_P, _I, _P, _I, _P, _I64, _I, _D, _U32, _P, _P],
    "mmmp_collide_segments": [_P, _I, _P, _P, _I, _I64, _I, _D, _U32, _P, _P],
    "mmmp_collide_robot_pairs": [_P, _I, _I, _P, _P, _I, _I64, _P, _P],
    "mmmp_compact_free": [_P, _I64, _P, _P, _P],
    "mmmp_gather_rows": [_P, _I, _P, _I64, _P, _P],
    "mmmp_knn": [_P, _P, _I64, _P, _P, _I64, _I, _I, _I, C.POINTER(Metric), _I, _I, _I, _I64, _P, _P, _P],
    "mmmp_knn_shard": [_P, _P, _I64, _I, _I, _I, C.POINTER(Metric), _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "mmmp_knn_range": [_P, _P, _I64, _I, _I, _I, C.POINTER(Metric), _I, _I, _D, _D, _P, _P, _P, _P, _P],
    "mmmp_pack_neighbour_records": [_P, _P, _P, _P, _I64, _I, _I64, _P, _P],
    "mmmp_scatter_neighbour_records": [_P, _I64, _I, _I64, _P, _P, _P, _P],
    "mmmp_knn_stats": [_I, _P],
    "mmmp_scan_order_cell": [_I, _I, _I, _I, _P],
    "mmmp_edge_stats": [_I, _P],
    "mmmp_knn_refine": [_P, _P, _I, C.POINTER(Metric), _P, _P, _I64, _I, _I, _D, _P, _P, _P, _P],
    "mmmp_radius": [_P, _P, _I64, _P, _P, _I64, _I, _I, C.POINTER(Metric), C.POINTER(Metric), _D, _I, _I, _I, _I,
                    _I64, _P, _P, _P, _P, _P],
    "mmmp_exclusive_scan_i32": [_P, _I64, _P, _P],
    "mmmp_edges_from_knn": [_P, _I64, _I, _P, _P],
    "mmmp_collide_knn_edges": [_P, _I, _P, _I, _P, _I64, _I, _I64, _I64, _I, _I, _D, _U32, _P, _P],
    "mmmp_resolve_knn_edges": [_P, _I64, _I, _I64, _I64, _P, _P],
    "mmmp_roadmap_candidates_host": [_P, _I, _P, _I64, _I, _I, _I, _I, _D, _U32, _P, _P, _P, _P, _P, _P],
    "mmmp_roadmap_build_host": [_P, _I, _P, _I64, _I, _I, _I, _I, _I, _D, _U32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "mmmp_greedy_degree_connect": [_P, _P, _I64, _I, _I, _I, _P, _P],
    "mmmp_greedy_degree_connect_device": [_P, _P, _I64, _I, _I, _I, _P, _P, _P, _P],
    "mmmp_version": [],
    "mmmp_device_info": [_P, _P, _P, _P],
    "mmmp_connected_components": [_P, _P, _I64, _I, _P, _P, _P],
    "mmmp_ik_dls": [_P, _I, _P, _P, _P, _I, _I64, _P, _P, _I, _D, _D, _D, _D, _P, _I, _P, _P, _P],
    "mmmp_launch_count": [],
}
_RESTYPES = {"mmmp_launch_count": C.c_int64}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmmmp_b200.so")
if os.environ.get("MMMP_B200_LIB"):    # kernel-tuning builds of the same library (tools/build_variant.py)
    LIB_PATH = os.path.abspath(os.environ["MMMP_B200_LIB"])
_lib = None


def lib_path() -> str:
    return LIB_PATH


def load():
    """Load libmmmp_b200.so (built in-tree by mmmp_b200._build); fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MmmpError(
            f"{LIB_PATH} is missing: build it with `python -m mmmp_b200._build` "
            "(there is no CPU fallback for the roadmap-construction path)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc >= 1000:
        raise MmmpError(f"{what}: CUDA error {rc - 1000} (no usable CUDA device or a kernel failed; "
                        "mmmp_b200 has no CPU fallback)")
    names = {1: "invalid argument", 2: "limit exceeded", 3: "wrong world kind / metric kind"}
    raise MmmpError(f"{what}: {names.get(rc, 'error')} (status {rc})")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
