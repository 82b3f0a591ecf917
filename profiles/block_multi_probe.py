"""Block CG on several GPUs (one rank per GPU under torchrun): 4 sources through the K-wide stencil with one halo exchange
for the 4 inputs, against 4 single solves on the same partitioned context.  BASELINE configs[3] lattice by default."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api, dist as D  # noqa: E402

EVEN = 2
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local_rank = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local_rank)
dist.init_process_group("gloo")
dims = tuple(int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (64, 64, 64, 96)
grid = D.rank_grid(world)
ids = [api.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx = api.Context(dims, device=local_rank, grid=grid, rank=rank, nranks=world, nccl_id=ids[0])
ctx.links_synthetic(1234)
vb = [ctx.vec_create() for _ in range(4)]
vx = [ctx.vec_create() for _ in range(4)]
for k in range(4):
    ctx.vec_gaussian(vb[k], EVEN, 5678 + k)
out = {"lattice": list(dims), "n_gpus": world, "rank_grid": list(grid)}
for mixed in (1, 0):
    for rep in range(2):
        for v in vx:
            ctx.vec_zero(v, EVEN)
        tot, res = ctx.congrad_block_dev(vb, vx, 0.05, EVEN, 2000, 10, 1e-10, mixed_precision=mixed)
    row = {"block_seconds": res[0]["device_seconds"], "block_iterations": tot, "converged": [r["converged"] for r in res]}
    s = 0.0
    its = 0
    for k in range(4):
        ctx.vec_zero(vx[k], EVEN)
        it, r = ctx.congrad_dev(vb[k], vx[k], 0.05, EVEN, 2000, 10, 1e-10, mixed_precision=mixed)
        s += r["device_seconds"]
        its += it
    row["four_single_solves_seconds"] = s
    row["single_iterations"] = its
    row["speedup"] = s / row["block_seconds"]
    V = 1
    for d in dims:
        V *= d
    row["block_gflops_milc_convention"] = 1187.0 * V * tot / row["block_seconds"] / 1e9   # (bench.py's convention)
    out["mixed_precision_%d" % mixed] = row
if rank == 0:
    print(json.dumps(out))
ctx.close()
dist.barrier()
dist.destroy_process_group()
