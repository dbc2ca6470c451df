"""Parity of the path bench.py TIMES ("pin the path you benchmark").

The neighbour scan takes its culled branch (tiles sorted along the Hilbert curve, bounding-box culling, the
fp32 centroid filter with its 2e-5 slack) only at >= 4096 references, so the golden vectors
of tests/test_gpu_kernels.py (<= 3000 rows) never reach it.  Here the culled scan, the k + 2
candidate pad, the fp64 re-rank and the re-search of ambiguous rows run at the BASELINE.json
sizes -- C2 (10 000 nodes, all rows), C4 (200 000 composite 14-D rows), C5 (1 000 000 nodes) --
through RoadmapBuilder.candidates, the call bench.py's step() makes, and are compared with a
float64 brute force:

  * positions / rows come from the ORACLE (oracle/chain.py: panda.py:111-173 restated; the
    14-D rows are the configurations themselves);
  * the distance of every (query, reference) pair is evaluated in float64 in the reference's
    order -- task_space_pose_distance, panda.py:248-254: sum over the 8 frames of the norm of
    the position difference; general_distance, coupled_prm.py:131 -- with plain torch float64
    arithmetic on the GPU (the checker; none of the library's kernels), because 2048 x 10^6
    pairs are too many for numpy inside a test;
  * bar: distances equal to 1e-10, indices equal except inside float64 ties < 1e-6
    (north_star check 1; decoupled_prm.py:480-494 heap-loop semantics).

Edge verdicts of the same builds are compared with the float64 sphere-model oracle on a
random sample of candidate edges.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LIMITS = np.array([(-2.8973, 2.8973), (-1.7628, 1.7628), (-2.8973, 2.8973), (-3.0718, -0.0698),
                   (-2.8973, 2.8973), (-0.0175, 3.7525), (-2.8973, 2.8973)])
DIST_TOL = 1e-10      # float64 distances (north_star: FK metric re-ranked in the reference's order)
TIE_TOL = 1e-6        # documented: indices may differ only inside distance ties below this


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def core(cuda):
    from mmmp_b200 import core
    return core


def bruteforce_topk(rows64, sel, k, grouped, chunk_bytes=4 << 30):
    """float64 brute force on the checker side.  rows64 [n, G, 3] (grouped: sum over the G
    groups of the 3-D norm, in group order) or [n, D] (plain L2) as a CUDA float64 tensor;
    sel: query rows.  -> (dist [m, k] ascending, idx [m, k] ordered by (dist, id), D_of(q, ids))."""
    n = rows64.shape[0]
    per_q = rows64[0].numel() * 8 * n * 3            # diff, squares, norms
    step = max(1, int(chunk_bytes // per_q))
    out_d, out_i = [], []
    kk = min(k + 8, n - 1)
    sel_t = torch.as_tensor(sel, device=rows64.device, dtype=torch.long)
    for a in range(0, len(sel), step):
        s = sel_t[a:a + step]
        diff = rows64[None, :, :] - rows64[s][:, None]
        if grouped:
            norms = diff.square().sum(-1).sqrt()                    # [m, n, G]
            d = torch.zeros(norms.shape[:2], dtype=torch.float64, device=rows64.device)
            for g in range(norms.shape[2]):                         # the reference's summation order
                d = d + norms[:, :, g]
        else:
            d = diff.square().sum(-1).sqrt()
        d[torch.arange(len(s), device=d.device), s] = float("inf")  # exclude self
        dv, di = torch.topk(d, kk, dim=1, largest=False, sorted=True)
        out_d.append(dv.cpu().numpy())
        out_i.append(di.cpu().numpy())
        del diff, d
    dv, di = np.concatenate(out_d), np.concatenate(out_i)
    order = np.lexsort((di, dv), axis=1)                            # ascending (dist, id)
    dv, di = np.take_along_axis(dv, order, 1), np.take_along_axis(di, order, 1)
    return dv[:, :k], di[:, :k], dv, di


def pair_distance(rows64, qi, ri, grouped):
    diff = rows64[torch.as_tensor(ri, device=rows64.device).long()] - rows64[torch.as_tensor(qi, device=rows64.device).long()]
    if grouped:
        norms = diff.square().sum(-1).sqrt()
        d = torch.zeros(norms.shape[0], dtype=torch.float64, device=rows64.device)
        for g in range(norms.shape[1]):
            d = d + norms[:, g]
        return d.cpu().numpy()
    return diff.square().sum(-1).sqrt().cpu().numpy()


def check_neighbours(rows64, sel, idx, dist, k, grouped):
    ref_d, ref_i, _, _ = bruteforce_topk(rows64, sel, k, grouped)
    got_d, got_i = dist[sel], idx[sel]
    assert np.abs(got_d - ref_d).max() < DIST_TOL, f"k smallest float64 distances differ by {np.abs(got_d - ref_d).max()}"
    assert (got_i != np.asarray(sel)[:, None]).all(), "a node came back as its own neighbour"
    # the ids carry those distances ...
    d_of_ids = pair_distance(rows64, np.repeat(sel, k), got_i.ravel(), grouped).reshape(got_i.shape)
    assert np.abs(d_of_ids - got_d).max() < DIST_TOL
    # ... and differ from (dist, id) order only inside ties below TIE_TOL
    diff = got_i != ref_i
    for r, c in zip(*np.nonzero(diff)):
        assert abs(ref_d[r, c] - d_of_ids[r, c]) < TIE_TOL, (r, c, ref_d[r, c], d_of_ids[r, c])
    for r in range(len(sel)):
        assert len(set(got_i[r].tolist())) == k, "duplicate neighbour id"
    return float(diff.mean())


def _c5_world(core, arm=0):
    model = core.load_model("panda_spheres")
    ang = arm * np.pi / 2
    yaw = ang + np.pi
    base = core.base_matrix((0.665 * np.cos(ang), 0.665 * np.sin(ang), 0.01), (0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2)))
    obstacles = [core.plane((0, 0, 1), 0.0)]
    return model, base, obstacles, core.SpatialWorld([core.RobotSpec(model, base)], obstacles)


@pytest.mark.parametrize("n,n_query", [(10_000, 10_000), (200_000, 2048), (1_000_000, 2048)])
def test_fk_metric_culled_scan_vs_float64_bruteforce(core, cuda, n, n_query):
    """C2 (all rows), and C5-sized builds (2048 random query rows): RoadmapBuilder.candidates
    with the FK metric = culled scan + k+2 pad + fp64 re-rank, the path bench.py times."""
    from mmmp_b200.roadmap import RoadmapBuilder
    from oracle import chain
    from oracle.spheres import SphereChecker
    model, base, obstacles, world = _c5_world(core)
    b = RoadmapBuilder(world, 0, "fk", 10, 50, 0.02, single_rank=True)
    q64, q32 = b.sample_free(n, seed=77)
    out = b.candidates(q64, q32)
    q = _np(q64)
    idx, dist, free = _np(out["idx"]), _np(out["dist"]), _np(out["edge_free"]).astype(bool)
    assert idx.shape == (n, 10) and (idx >= 0).all() and (idx < n).all()
    assert int(_np(out["ambiguous"]).sum()) == 0, "rows left ambiguous after the wide re-search"
    rng = np.random.default_rng(n)
    sel = np.arange(n) if n_query >= n else np.sort(rng.choice(n, n_query, replace=False))
    P = torch.from_numpy(np.ascontiguousarray(chain.panda_positions(q, base))).to(cuda)   # oracle positions [n, 8, 3]
    frac = check_neighbours(P, sel, idx, dist, 10, grouped=True)
    assert frac < 0.01
    del P
    # every sample is collision free under the float64 sphere-model oracle, and so say the edge verdicts
    ck = SphereChecker([model], [base], obstacles)
    some = rng.choice(n, 2000, replace=False)
    assert ck.config_collision(q[some], 0).sum() <= 2
    rows = rng.choice(n, 100, replace=False)
    eref = ~ck.edge_collision(np.repeat(q[rows], 10, 0), q[idx[rows].ravel()], 50, 0.02, 0)
    assert (free[rows].ravel() != eref).sum() <= 3
    world.close()


def test_composite_14d_scan_vs_float64_bruteforce(core, cuda):
    """C4: 2 Pandas coupled, 200 000 composite 14-D rows, joint-space L2 (coupled_prm.py:131),
    the 16-column instantiation of the scan with tile culling."""
    from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_ROBOTS, CHECK_SELF
    from mmmp_b200.roadmap import RoadmapBuilder
    from oracle.spheres import SphereChecker
    model = core.load_model("panda_spheres")
    bases = [core.base_matrix((0.7, 0, 0.01), (0, 0, 1, 0)), core.base_matrix((-0.7, 0, 0.01))]
    obstacles = [core.plane((0, 0, 1), 0.0)]
    world = core.SpatialWorld([core.RobotSpec(model, b, 7 * i) for i, b in enumerate(bases)], obstacles)
    F = CHECK_SELF | CHECK_OBSTACLES | CHECK_ROBOTS
    n = 200_000
    b = RoadmapBuilder(world, -1, "l2", 10, 20, 0.05, flags=F, single_rank=True)
    q64, q32 = b.sample_free(n, seed=5)
    out = b.candidates(q64, q32)
    idx, dist, free = _np(out["idx"]), _np(out["dist"]), _np(out["edge_free"]).astype(bool)
    rng = np.random.default_rng(14)
    sel = np.sort(rng.choice(n, 2048, replace=False))
    frac = check_neighbours(q64, sel, idx, dist, 10, grouped=False)
    assert frac < 0.02          # the 0.01 lattice makes exact L2 ties common
    q = _np(q64)
    ck = SphereChecker([model, model], bases, obstacles)
    some = rng.choice(n, 500, replace=False)
    assert ck.composite_collision(q[some]).sum() <= 2
    rows = rng.choice(n, 60, replace=False)
    qa, qb = np.repeat(q[rows], 10, 0), q[idx[rows].ravel()]
    hit = np.zeros(len(qa), bool)
    for i in range(20):
        hit |= ck.composite_collision(qa + (i * 0.05) * (qb - qa))
    assert (free[rows].ravel() != ~hit).sum() <= 3
    world.close()


def test_l2_7d_culled_scan_vs_float64_bruteforce(core, cuda):
    """Joint-space L2 of one arm at 200 000 nodes (general_distance, decoupled_prm.py:824): the
    8-column instantiation with culling."""
    from mmmp_b200.roadmap import RoadmapBuilder
    model, base, obstacles, world = _c5_world(core, arm=1)
    n = 200_000
    b = RoadmapBuilder(world, 0, "l2", 10, 50, 0.02, single_rank=True)
    q64, q32 = b.sample_free(n, seed=9)
    out = b.candidates(q64, q32)
    sel = np.sort(np.random.default_rng(2).choice(n, 2048, replace=False))
    frac = check_neighbours(q64, sel, _np(out["idx"]), _np(out["dist"]), 10, grouped=False)
    assert frac < 0.02
    world.close()


def test_sharded_build_equals_single_rank_build(core, cuda):
    """The multi-GPU decomposition on ONE device: the slices mmmp_knn_range cuts for 2, 4 and 8
    ranks (with the uneven cuts the balancer produces), re-ranked and scattered back the way
    RoadmapBuilder.candidates does, reproduce the single-rank tables of a 300 000-node build bit
    for bit -- culled path.  (bench.py repeats this check across real ranks: "shard_parity".)"""
    from mmmp_b200.roadmap import RoadmapBuilder
    model, base, obstacles, world = _c5_world(core, arm=2)
    n = 300_000
    b = RoadmapBuilder(world, 0, "fk", 10, 50, 0.02, single_rank=True)
    q64, q32 = b.sample_free(n, seed=31)
    ref = b.candidates(q64, q32)
    feats = b.features(q64, q32)
    for cuts in (np.linspace(0, 1, 3), np.array([0.0, 0.21, 0.48, 0.80, 1.0]), np.linspace(0, 1, 9)):
        idx = torch.full((n, 10), -7, dtype=torch.int32, device=cuda)
        dist = torch.full((n, 10), -1.0, dtype=torch.float64, device=cuda)
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            rows, ci, cd = core.knn_range(feats, 12, float(lo), float(hi))
            i2, d2, amb = core.knn_refine(feats.x64, core.gather_rows(feats.x64, rows), feats.m64, ci, cd, 10)
            assert int(amb.sum()) == 0
            idx[rows.long()], dist[rows.long()] = i2, d2
        assert torch.equal(idx, ref["idx"]) and torch.equal(dist, ref["dist"])
    world.close()


def test_kuka_roadmap_build_vs_float64_oracle(core, cuda):
    """The KUKA iiwa14 chain (URDF joint origins, the reference's own collision spheres) through the
    same builder: samples free of the obstacles, the FK metric over the frames of ITS chain
    (mmmp_fk_metric_groups on general origins: no modified-DH shortcut), candidates, edge verdicts,
    accept pass.  Neighbours against a float64 brute force on the oracle's generic chain
    (oracle/chain.chain_frames), edges against the float64 sphere oracle.  (Self collision is off: the
    URDF's spheres of links two apart always overlap, geometry/kuka_iiwa14_spheres.json.)"""
    from mmmp_b200._lib import CHECK_OBSTACLES
    from mmmp_b200.roadmap import RoadmapBuilder
    from oracle import chain
    from oracle.spheres import SphereChecker, model_origins
    m = core.load_model("kuka_iiwa14_spheres")
    base = core.base_matrix((0.1, -0.2, 0.0), (0, 0, 0.38268343, 0.92387953))
    obstacles = [core.plane((0, 0, 1), -0.2), core.box((0.5, 0.1, 0.5), (0.1, 0.3, 0.1)), core.sphere((-0.3, 0.4, 0.8), 0.2)]
    world = core.SpatialWorld([core.RobotSpec(m, base)], obstacles)
    n, k = 60_000, 8
    b = RoadmapBuilder(world, 0, "fk", k, 20, 0.05, flags=CHECK_OBSTACLES, single_rank=True)
    out = b.build(n, seed=12, k2=6)
    q = _np(out["q64"])
    idx, dist, free = _np(out["idx"]), _np(out["dist"]), _np(out["edge_free"]).astype(bool)
    ck = SphereChecker([m], [base], obstacles)
    rng = np.random.default_rng(6)
    some = rng.choice(n, 1500, replace=False)
    assert ck.config_collision(q[some], 0, self_=False).sum() <= 2
    # FK metric of the reference's definition on this chain: sum over all frames of the origin distances
    P = chain.chain_frames(model_origins(m), q, base)[:, :, :3, 3]            # [n, 8, 3], float64 oracle
    sel = np.sort(rng.choice(n, 1024, replace=False))
    frac = check_neighbours(torch.from_numpy(np.ascontiguousarray(P)).to(cuda), sel, idx, dist, k, grouped=True)
    assert frac < 0.01
    rows = rng.choice(n, 80, replace=False)
    eref = ~ck.edge_collision(np.repeat(q[rows], k, 0), q[idx[rows].ravel()], 20, 0.05, 0, self_=False)
    assert (free[rows].ravel() != eref).sum() <= 3
    # the accepted adjacency is the sequential pass on these tables
    adj, deg = core.greedy_degree_connect(idx, free.astype(np.uint8), 6)
    assert (_np(out["adj"]) == adj).all() and (_np(out["deg"]) == deg).all()
    world.close()
