"""CPU suite: the C-ABI library loads, exports every symbol include/mmmp_b200.h declares,
the ctypes mirrors have the C struct sizes, host-only entry points work, and compute
entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from mmmp_b200 import _build, _lib, core

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mmmp_b200.h")


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(mmmp_[a-z0-9_]+)\s*\(", text))
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"


def test_struct_sizes_match_c(lib):
    src = r'''
#include <stdio.h>
#include "mmmp_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(mmmp_robot_desc), sizeof(mmmp_obstacle_desc), sizeof(mmmp_planar_desc),
         sizeof(mmmp_metric), sizeof(mmmp_pair_table));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o",
                               os.path.join(d, "s")])
        out = subprocess.check_output([os.path.join(d, "s")]).split()
    sizes = [int(x) for x in out]
    assert sizes == [C.sizeof(_lib.RobotDesc), C.sizeof(_lib.ObstacleDesc), C.sizeof(_lib.PlanarDesc),
                     C.sizeof(_lib.Metric), C.sizeof(_lib.PairTable)]


def test_greedy_connect_matches_oracle(lib, golden):
    from oracle import knn as oknn
    g = golden("panda_heap_knn")
    qs = g["samples"]
    idx = g["l2_knn_idx"].astype(np.int32)
    def ef(i, j):
        a, b = qs[i], qs[j]
        return (int(round((sum(a) + sum(b)) * 100)) % 7) != 0
    free = np.array([[ef(i, j) for j in row] for i, row in enumerate(idx)], dtype=np.uint8)
    adj, deg = core.greedy_degree_connect(idx, free, k2=6)
    ref_deg, ref_adj = g["l2_degree_deg"], g["l2_degree_adj"]
    assert (deg == ref_deg).all()
    flat = np.concatenate([adj[i, :deg[i]] for i in range(len(deg))])
    assert (flat == ref_adj).all()


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_scan_order_is_a_hilbert_curve(lib):
    """The neighbour search sorts its rows by grid cell along mmmp_scan_order_cell (host function).
    What the tile culling relies on: the numbering is a bijection of the grid and consecutive cells
    share a face, so a run of consecutive rows never jumps across the grid."""
    for bits in (1, 2, 3, 4):
        n = 1 << bits
        inv = np.full((n ** 3, 3), -1, np.int64)
        out = C.c_uint32(0)
        for x in range(n):
            for y in range(n):
                for z in range(n):
                    assert lib.mmmp_scan_order_cell(x, y, z, bits, C.byref(out)) == 0
                    assert out.value < n ** 3 and inv[out.value, 0] < 0, "not a bijection"
                    inv[out.value] = (x, y, z)
        steps = np.abs(np.diff(inv, axis=0)).sum(1)
        assert (steps == 1).all(), f"bits={bits}: {int((steps != 1).sum())} steps between non-adjacent cells"
    # spot checks on the 6-bit grid the 1 M-row search uses, and argument errors
    rng = np.random.default_rng(0)
    seen = set()
    for x, y, z in rng.integers(0, 64, (2000, 3)):
        assert lib.mmmp_scan_order_cell(int(x), int(y), int(z), 6, C.byref(out)) == 0
        seen.add((out.value, (int(x), int(y), int(z))))
    assert len({i for i, _ in seen}) == len({c for _, c in seen})
    assert lib.mmmp_scan_order_cell(64, 0, 0, 6, C.byref(out)) != 0
    assert lib.mmmp_scan_order_cell(0, 0, 0, 0, C.byref(out)) != 0
    assert lib.mmmp_scan_order_cell(0, 0, 0, 3, None) != 0


def test_no_cpu_fallback(lib):
    with pytest.raises(_lib.MmmpError):
        core.sample_uniform([0.0], [1.0], 4, seed=1)
    d = _lib.PlanarDesc()
    d.n_arms = 1
    d.n_links[0] = 2
    h = C.c_void_p(0)
    rc = lib.mmmp_planar_world_create(C.byref(d), C.byref(h))
    assert rc >= 1000        # a CUDA error, not a silent CPU path


def test_product_package_never_imports_oracle():
    import ast
    pkg = os.path.join(ROOT, "mmmp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom) and node.module:
                    mods = [node.module]
                assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), (dirpath, f)
