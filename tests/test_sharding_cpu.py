"""CPU suite, world_size 2 over gloo: the host-side logic of the multi-GPU path
(mmmp_b200/roadmap.py) -- contiguous row sharding, rank-ordered gather of ragged row blocks,
and the counter-based sampling blocks that make the sharded node order identical to the
single-GPU order (checked with the oracle restatement of the Philox sampler)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from mmmp_b200.roadmap import all_gather_rows, shard_rows
        from oracle import sampling
        ok = True
        # ragged gather: rank r contributes r + 2 rows
        t = torch.full((rank + 2, 3), float(rank)) + torch.arange(rank + 2)[:, None]
        g = all_gather_rows(t, size)
        want = torch.cat([torch.full((r + 2, 3), float(r)) + torch.arange(r + 2)[:, None] for r in range(size)])
        ok &= bool((g == want).all())
        # equal blocks and an empty block
        e = all_gather_rows(torch.zeros((0, 2)) if rank == 1 else torch.ones((4, 2)), size)
        ok &= e.shape == (4, 2)
        # row sharding covers [0, n) exactly once, in rank order
        for n in (1, 7, 1000, 1001):
            lo, hi = shard_rows(n, rank, size)
            blocks = all_gather_rows(torch.arange(lo, hi)[:, None], size)
            ok &= bool((blocks[:, 0] == torch.arange(n)).all())
        # spatial shards: every rank answers an arbitrary subset of the rows (mmmp_knn_shard);
        # gather of (row ids, values) + scatter restores the caller's row order
        n = 1003
        perm = torch.from_numpy(np.random.default_rng(5).permutation(n))
        lo, hi = shard_rows(n, rank, size)
        rows = perm[lo:hi]
        vals = (rows * 10).to(torch.float64)[:, None].repeat(1, 2)
        at = all_gather_rows(rows, size).long()
        part = all_gather_rows(vals, size)
        full = torch.empty((n, 2), dtype=part.dtype)
        full[at] = part
        ok &= bool((full[:, 0] == torch.arange(n) * 10).all())
        # the same through the padded single-collective form RoadmapBuilder.candidates uses:
        # padding rows carry the sentinel id n and land in a spare row
        from mmmp_b200.roadmap import all_gather_padded
        cap = (n + size - 1) // size + 7
        at = all_gather_padded(rows, cap, size, fill=n).long()
        part = all_gather_padded(vals, cap, size)
        full = torch.empty((n + 1, 2), dtype=part.dtype)
        full[at] = part
        ok &= bool((full[:n, 0] == torch.arange(n) * 10).all()) and at.shape[0] == size * cap
        # sampling blocks: rank r draws candidates [first + r*per, first + (r+1)*per); the rank-ordered
        # concatenation equals the single-rank stream (counter-based generator)
        lo_, hi_ = np.array([-2.0, 0.0, -1.0]), np.array([2.0, 3.0, 1.0])
        per, first, seed = 500, 12345, 99
        mine = sampling.sample_uniform(lo_, hi_, per, seed, first + rank * per)
        free = mine[mine[:, 0] > -1.0]                       # a stand-in for the collision filter
        allf = all_gather_rows(torch.from_numpy(free), size).numpy()
        single = sampling.sample_uniform(lo_, hi_, per * size, seed, first)
        ok &= bool((allf == single[single[:, 0] > -1.0]).all())
        # arm-parallel layout: 4 arms on 2 ranks -> rank r builds arms r and r + 2 alone; the
        # hand-over leaves every arm's tables on rank 0
        from mmmp_b200.roadmap import ArmTeams
        teams = ArmTeams(4)
        ok &= teams.my_arms == [rank, rank + 2] and teams.team(rank) == (None, True) and teams.leader(3) == 1
        mine = {a: [torch.full((5, 3), a, dtype=torch.int32), torch.arange(5, dtype=torch.int32) + 10 * a]
                for a in teams.my_arms}
        like = {a: [((5, 3), torch.int32), ((5,), torch.int32)] for a in range(4)}
        got = teams.collect_on_root(mine, like)
        if rank == 0:
            ok &= sorted(got) == [0, 1, 2, 3]
            ok &= all(bool((got[a][0] == a).all()) and bool((got[a][1] == torch.arange(5) + 10 * a).all()) for a in range(4))
        # one arm on 2 ranks: the team is the whole world (sharded build inside the team)
        one = ArmTeams(1)
        ok &= one.ranks_of_arm == [[0, 1]] and one.team(0) == (None, False) and one.my_arms == [0]
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_sharding_logic_world_size_2():
    size = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(size, _free_port(), out), nprocs=size, join=True)
    assert all(out.get(r) for r in range(size)), dict(out)


def test_cut_controller_balances_a_skewed_scan_order():
    """roadmap.balanced_cuts (host logic of the sharded neighbour search): with a cost per query
    block that differs between regions of the scan order, repeated steps equalise the ranks'
    times; equal times, bad times and a single rank leave the cuts alone; the cuts stay a
    monotone partition of [0, 1] (what mmmp_knn_range expects)."""
    from mmmp_b200.roadmap import balanced_cuts

    def cost(lo, hi):                       # integral of a density 1 + 0.6 x over [lo, hi)
        return (hi - lo) + 0.3 * (hi * hi - lo * lo)

    for g in (2, 3, 4):
        cuts = np.linspace(0.0, 1.0, g + 1)
        first = np.array([cost(a, b) for a, b in zip(cuts[:-1], cuts[1:])])
        for _ in range(12):
            times = np.array([cost(a, b) for a, b in zip(cuts[:-1], cuts[1:])])
            cuts = balanced_cuts(cuts, times)
            assert cuts[0] == 0.0 and cuts[-1] == 1.0 and np.all(np.diff(cuts) > 0)
        times = np.array([cost(a, b) for a, b in zip(cuts[:-1], cuts[1:])])
        assert times.max() / times.min() < 1.01 < first.max() / first.min()
        assert times.max() < first.max()
    even = np.linspace(0.0, 1.0, 5)
    assert np.allclose(balanced_cuts(even, np.full(4, 7.5)), even)
    assert np.array_equal(balanced_cuts(even, np.array([1.0, 0.0, 2.0, 1.0])), even)      # a rank without a time
    assert np.array_equal(balanced_cuts(even, np.array([1.0, 2.0])), even)                # wrong length
    assert np.array_equal(balanced_cuts(np.array([0.0, 1.0]), np.array([3.0])), np.array([0.0, 1.0]))
