#!/usr/bin/env python
"""Multi-GPU consistency check (run under torchrun on the GPU box):
the roadmap built by G ranks (sample blocks + all-gather, sharded k-NN / edge checks +
all-gather) must equal, bit for bit, the roadmap one rank builds alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tools/check_multigpu.py [n_nodes]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import scene  # noqa: E402
from mmmp_b200 import core  # noqa: E402
from mmmp_b200.roadmap import RoadmapBuilder  # noqa: E402

rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
bases, obstacles = scene(4)
world = core.SpatialWorld([core.RobotSpec(core.load_model("panda_spheres"), b, 7 * i) for i, b in enumerate(bases)],
                          obstacles)
ok = True
for arm in (0, 3):
    sharded = RoadmapBuilder(world, arm, "fk", 10).build(n, seed=11 + arm)
    alone = RoadmapBuilder(world, arm, "fk", 10, single_rank=True).build(n, seed=11 + arm)
    for key in ("q64", "idx", "dist", "edge_free"):
        same = bool((sharded[key] == alone[key]).all())
        ok &= same
        if rank == 0:
            print(f"arm {arm} {key:9s} identical across sharding: {same}  shape {tuple(sharded[key].shape)}")
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI-GPU CONSISTENCY:", "PASS" if flag.item() == 1 else "FAIL", f"(world size {dist.get_world_size()})")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
