"""Roadmap build up to the adjacency on the device: the degree-capped greedy
accept pass of connect_nearest_neighbors_degree (decoupled_prm.py:496-503) as data-parallel
rounds (csrc/greedy.cu), against

  * the oracle's sequential restatement oracle/knn.greedy_degree (pinned by the reference's own
    production method run unbound, tests/golden/panda_heap_knn.npz and the 12-node known answer),
  * the native host replay mmmp_greedy_degree_connect (already pinned the same way),

bit for bit: same edges AND same order inside every adjacency list (the order decides what the
Python planners' edge_dicts lists look like).  Then the whole build of one C5 arm -- candidates ->
accept pass -> connected components -- at 1 M nodes against the host replay and a host union-find.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def core(cuda):
    from mmmp_b200 import core
    return core


def _oracle_adj(idx, free, k2, cap):
    """oracle/knn.greedy_degree restated on a candidate TABLE (holes, self references and
    duplicates inside a row are visited in turn, as the reference's loop would visit a list)."""
    n, k1 = idx.shape
    adj = [[] for _ in range(n)]
    for i in range(n):
        for c in range(k1):
            j = int(idx[i, c])
            if len(adj[i]) == k2:
                break
            if j < 0 or j >= n or j == i:
                continue
            if free[i, c] and j not in adj[i] and len(adj[j]) != k2:
                adj[i].append(j)
                adj[j].append(i)
    out = np.full((n, cap), -1, np.int32)
    for i, a in enumerate(adj):
        out[i, :len(a)] = a
    return out


@pytest.mark.parametrize("n,k1,k2,pfree", [(1, 3, 2, 1.0), (2, 1, 1, 1.0), (7, 4, 2, 0.8), (500, 10, 10, 0.75),
                                           (500, 10, 3, 1.0), (3000, 16, 6, 0.5), (3000, 5, 1, 0.9),
                                           (20000, 10, 10, 0.75), (20000, 10, 4, 1.0)])
def test_greedy_rounds_equal_sequential_pass(core, cuda, n, k1, k2, pfree):
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(n * 31 + k1 * 7 + k2)
    X = rng.random((n, 4))
    kk = min(k1 + 1, n)
    _, nb = cKDTree(X).query(X, k=kk)
    nb = nb.reshape(n, kk)
    idx = np.full((n, k1), -1, np.int32)
    idx[:, :kk - 1] = nb[:, 1:]
    free = (rng.random((n, k1)) < pfree).astype(np.uint8)
    if n >= 500:      # hostile rows: duplicates, self references, holes, out-of-range ids
        r = rng.choice(n, 40, replace=False)
        idx[r[:10], 3] = idx[r[:10], 1]
        idx[r[10:20], 2] = r[10:20]
        idx[r[20:30], 0] = -1
        idx[r[30:], 1] = n + 5
    cap = k2 + 2
    adj, deg, rounds = core.greedy_degree_connect_device(torch.from_numpy(idx).to(cuda), torch.from_numpy(free).to(cuda),
                                                         k2, cap, return_rounds=True)
    adj, deg = _np(adj), _np(deg)
    ref = _oracle_adj(idx, free, k2, cap)
    assert (adj == ref).all(), f"adjacency differs from the sequential pass in {(adj != ref).any(1).sum()} rows"
    assert (deg == (ref >= 0).sum(1)).all() and deg.max(initial=0) <= k2
    host_adj, host_deg = core.greedy_degree_connect(idx, free, k2, cap)
    assert (host_adj == adj).all() and (host_deg == deg).all()
    assert 1 <= rounds < 200


def test_greedy_rounds_on_a_dependency_chain(core, cuda):
    """Worst case for the round structure: nodes on a line, candidates = the two line neighbours,
    k2 = 1.  Node 0 takes node 1; node 1 is then full; node 2 must wait for that to know that it
    may take node 3; ... every decision hangs on the previous one, so the pass needs thousands of
    rounds -- and must still end with the sequential result (the round loop is unbounded)."""
    n = 20000
    idx = np.stack([np.arange(n) - 1, np.arange(n) + 1], 1).astype(np.int32)
    idx[idx >= n] = -1
    free = np.ones_like(idx, dtype=np.uint8)
    adj, deg, rounds = core.greedy_degree_connect_device(torch.from_numpy(idx).to(cuda), torch.from_numpy(free).to(cuda),
                                                         1, 1, return_rounds=True)
    ref = _oracle_adj(idx, free, 1, 1)
    assert (_np(adj) == ref).all() and rounds > 4096
    assert (_np(deg) == 1).all() and (_np(adj)[0::2, 0] == np.arange(1, n, 2)).all()


def test_greedy_matches_reference_golden(core, golden, cuda):
    """The reference's own connect_nearest_neighbors_degree output (production method run unbound
    on 300 Panda samples, tools/gen_golden.py): candidates from the golden k-NN table, all
    edges free."""
    g = golden("panda_heap_knn")
    idx = g["l2_knn_idx"].astype(np.int32)
    free = np.ones_like(idx, dtype=np.uint8)
    for k2 in (3, 6, 10):
        adj, deg = core.greedy_degree_connect_device(torch.from_numpy(idx).to(cuda), torch.from_numpy(free).to(cuda), k2)
        assert (_np(adj) == _oracle_adj(idx, free, k2, k2)).all()


def test_build_to_adjacency_1m(core, cuda):
    """One C5 arm, 1 M nodes: candidates -> accept pass -> components, all on the device; the
    adjacency equals the native host replay of the reference loop and the component labels
    equal a host union-find over the accepted edges."""
    import time
    from mmmp_b200.roadmap import RoadmapBuilder
    model = core.load_model("panda_spheres")
    base = core.base_matrix((0.665, 0, 0.01), (0, 0, 1, 0))
    world = core.SpatialWorld([core.RobotSpec(model, base)], [core.plane((0, 0, 1), 0.0)])
    n = 1_000_000
    b = RoadmapBuilder(world, 0, "fk", 10, 50, 0.02, single_rank=True)
    out = b.build(n, seed=3)
    for k2 in (10, 6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        adj, deg, rounds = core.greedy_degree_connect_device(out["idx"], out["edge_free"], k2, return_rounds=True)
        e1.record()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        host_adj, host_deg = core.greedy_degree_connect(_np(out["idx"]), _np(out["edge_free"]), k2)
        host_ms = (time.perf_counter() - t0) * 1e3
        print(f"accept pass 1M x 10, k2={k2}: device {e0.elapsed_time(e1):.2f} ms in {rounds} rounds, "
              f"host replay {host_ms:.1f} ms, mean degree {float(deg.float().mean()):.3f}")
        assert (_np(adj) == host_adj).all() and (_np(deg) == host_deg).all()
    labels, count = core.connected_components(adj)
    # host union-find over the accepted edges
    a = _np(adj)
    src = np.repeat(np.arange(n), a.shape[1])[a.ravel() >= 0]
    dst = a.ravel()[a.ravel() >= 0]
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    ncomp, lab = connected_components(coo_matrix((np.ones(len(src), np.int8), (src, dst)), shape=(n, n)), directed=False)
    assert ncomp == count
    first = np.full(ncomp, n, np.int64)
    np.minimum.at(first, lab, np.arange(n))
    assert (_np(labels) == first[lab]).all()
    world.close()


def test_roadmap_build_host_equals_device_pipeline(core, cuda):
    """mmmp_roadmap_build_host (the reference-facing call of bench.py's e2e figure: host samples in,
    the ROADMAP out, copies overlapped with kernels) against the device-resident pipeline pieces on
    the same samples: accept test, candidate tables, verdicts, accepted adjacency, components --
    with page-locked and with ordinary host buffers."""
    from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_SELF
    from mmmp_b200.roadmap import RoadmapBuilder
    model = core.load_model("panda_spheres")
    base = core.base_matrix((0.665, 0, 0.01), (0, 0, 1, 0))
    world = core.SpatialWorld([core.RobotSpec(model, base)], [core.plane((0, 0, 1), 0.0)])
    lim = np.array(model["joint_limits"])
    N, k1, k2 = 60_000, 10, 6
    q64, q32 = core.sample_uniform(lim[:, 0], lim[:, 1], N, seed=77)
    q_host = _np(q64)
    # device pipeline on the same samples
    hit = world.collide_configs(q32, 0, CHECK_SELF | CHECK_OBSTACLES)
    keep = core.compact_free(hit)
    n64 = core.gather_rows(q64, keep)
    b = RoadmapBuilder(world, 0, "fk", k1, 50, 0.02, single_rank=True)
    out = b.candidates(n64, core.pack_f32(n64))
    b.adjacency(out, k2)
    n = n64.shape[0]
    for pinned in (True, False):
        bufs = core.host_buffers(N, k1, k2, pinned=pinned)
        res = world.roadmap_build_host(q_host, 0, core.METRIC_GROUP_L2_SUM, k1, k2, 50, 0.02, out=bufs)
        assert res["n_nodes"] == n and (res["node_src"] == _np(keep)).all()
        assert (res["nbr_idx"] == _np(out["idx"])).all() and (res["nbr_dist"] == _np(out["dist"])).all()
        assert (res["edge_free"] == _np(out["edge_free"])).all()
        assert (res["adj"] == _np(out["adj"])).all() and (res["deg"] == _np(out["deg"])).all()
        assert (res["labels"] == _np(out["labels"])).all() and res["n_components"] == int(out["n_components"].item())
        t = res["timings_ms"]
        assert t[7] > 0 and t[8] > 0 and t[9] > 0 and abs(t[:6].sum() + t[8] + t[9] + t[6] - t[7]) < 0.2 * t[7] + 0.5
    # the candidates-only entry point returns the same tables
    res2 = world.roadmap_candidates_host(q_host, 0, core.METRIC_GROUP_L2_SUM, k1, 50, 0.02)
    assert (res2["nbr_idx"] == _np(out["idx"])).all() and (res2["edge_free"] == _np(out["edge_free"])).all()
    world.close()


def test_candidate_table_verdicts_share_mutual_pairs(core, cuda):
    """mmmp_collide_knn_edges: a pair of nodes that name each other is evaluated once, low id -> high id (the
    first of the two checks of decoupled_prm.py:496-503); every other slot row -> neighbour.  The table equals
    mmmp_collide_edges on exactly those directed edges bit for bit, however the rows are cut across callers,
    and nothing is shared when the t grid is not symmetric (n_steps * step != 1)."""
    import os
    from mmmp_b200._lib import CHECK_OBSTACLES, CHECK_SELF
    from mmmp_b200.roadmap import RoadmapBuilder
    F = CHECK_SELF | CHECK_OBSTACLES
    model = core.load_model("panda_spheres")
    world = core.SpatialWorld([core.RobotSpec(model, core.base_matrix((0, 0, 0.01)))],
                              [core.plane((0, 0, 1), 0.0), core.box((0.45, 0.0, 0.3), (0.08, 0.3, 0.3))])
    n, k = 30_000, 10
    b = RoadmapBuilder(world, 0, "fk", k, 50, 0.02, single_rank=True)
    q64, q32 = b.sample_free(n, seed=3)
    feats = b.features(q64, q32)
    idx, _, _ = core.knn_exact(feats, feats, k)
    idx[17, 4] = -1                                       # an empty slot
    idx[23, 0] = n + 5                                    # an id outside the table: treated like an empty slot
    nbr = _np(idx).astype(np.int64)
    nbr[nbr >= n] = -1
    rows = np.repeat(np.arange(n), k).reshape(n, k)
    valid = nbr >= 0
    mutual = np.zeros((n, k), bool)
    mutual[valid] = (nbr[nbr[valid]] == rows[valid][:, None]).any(1)
    assert 0.2 < mutual.mean() < 0.9
    a = np.where(mutual, np.minimum(rows, nbr), rows)
    c = np.where(mutual, np.maximum(rows, nbr), nbr)

    def directed(a, c, n_steps, step):
        e = torch.from_numpy(np.stack([a.ravel(), c.ravel()], 1).astype(np.int32)).to(cuda)
        return (_np(world.collide_edges(q32, e, n_steps, step, 0, F)) == 0).reshape(n, k).astype(np.uint8)

    want = directed(a, c, 50, 0.02)
    assert 0.03 < 1.0 - want.mean() < 0.9
    got = _np(world.collide_knn_edges(q32, idx, 50, 0.02, 0, F))
    assert (got == want).all() and got[17, 4] == 0 and got[23, 0] == 0
    # both slots of a mutual pair agree by construction
    j = nbr[mutual]
    s_back = (nbr[j] == rows[mutual][:, None]).argmax(1)
    assert (got[mutual] == got[j, s_back]).all()
    # row blocks of several callers: partners inside the block only ...
    cuts = [0, 9_000, 9_001, 22_222, n]
    parts = [_np(world.collide_knn_edges(q32, idx, 50, 0.02, 0, F, rows=(lo, hi))) for lo, hi in zip(cuts[:-1], cuts[1:])]
    assert (np.concatenate(parts) == want).all()
    # ... or anywhere, resolved on the gathered table; every pair is evaluated by exactly one block
    parts = [world.collide_knn_edges(q32, idx, 50, 0.02, 0, F, rows=(lo, hi), share_across_rows=True)
             for lo, hi in zip(cuts[:-1], cuts[1:])]
    table = torch.cat(parts).contiguous()
    marked = _np(table) == 2
    assert marked.any() and not (marked & ~mutual).any()
    assert (marked[mutual] != marked[j, s_back]).all(), "exactly one slot of a mutual pair is evaluated"
    core.SpatialWorld.resolve_knn_edges(idx, table)
    assert (_np(table) == want).all()
    # a grid that is not symmetric shares nothing: every slot row -> neighbour
    for n_steps, step in ((7, 0.15), (20, 0.04)):
        assert (_np(world.collide_knn_edges(q32, idx, n_steps, step, 0, F)) == directed(rows, nbr, n_steps, step)).all()
    # and the shared evaluation differs from two separate ones by rounding at most (same configurations)
    os.environ["MMMP_EDGES_NO_SHARE"] = "1"
    try:
        plain = _np(world.collide_knn_edges(q32, idx, 50, 0.02, 0, F))
    finally:
        del os.environ["MMMP_EDGES_NO_SHARE"]
    assert (plain == directed(rows, nbr, 50, 0.02)).all()
    assert (plain != want).sum() <= 5
    world.close()


def test_roadmap_build_host_from_concurrent_threads(core, cuda):
    """The host-buffer call creates its own streams and keeps no shared mutable state: two arms built
    from two host threads at once (what bench.py's e2e does) return what the calls return one by one."""
    from concurrent.futures import ThreadPoolExecutor
    model = core.load_model("panda_spheres")
    bases = [core.base_matrix((0.5, 0, 0.01), (0, 0, 1, 0)), core.base_matrix((-0.5, 0, 0.01))]
    world = core.SpatialWorld([core.RobotSpec(model, b, 7 * i) for i, b in enumerate(bases)], [core.plane((0, 0, 1), 0.0)])
    lim = np.array(model["joint_limits"])
    rng = np.random.default_rng(8)
    qs = [np.round(rng.uniform(lim[:, 0], lim[:, 1], (30_000, 7)), 2) for _ in range(2)]
    keys = ("node_src", "nbr_idx", "nbr_dist", "edge_free", "adj", "deg", "labels")

    dev_index = torch.cuda.current_device()

    def call(a):
        torch.cuda.set_device(dev_index)        # the current device is per host thread
        r = world.roadmap_build_host(qs[a], a, core.METRIC_GROUP_L2_SUM, 10, 6, 50, 0.02)
        return {k: np.array(r[k]) for k in keys} | {"n_nodes": r["n_nodes"], "n_components": r["n_components"]}

    one_by_one = [call(0), call(1)]
    with ThreadPoolExecutor(max_workers=2) as pool:
        for _ in range(3):
            together = list(pool.map(call, range(2)))
            for a in range(2):
                assert together[a]["n_nodes"] == one_by_one[a]["n_nodes"] > 20_000
                assert together[a]["n_components"] == one_by_one[a]["n_components"]
                for k in keys:
                    assert (together[a][k] == one_by_one[a][k]).all(), (a, k)
    world.close()
