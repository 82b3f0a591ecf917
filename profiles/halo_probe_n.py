"""N-GPU probe (torchrun): stencil and CG iteration time of BASELINE configs[3] (64^3x96 split
over N ranks) for the halo implementations and push-kernel widths.  Max over ranks."""
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
dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local_rank))
dims = (64, 64, 64, 96)
grid = D.rank_grid(world)
variants = [("p2p", c) for c in os.environ.get("PROBE_CTAS", "148,32").split(",")] + [("nccl", "0")]
out = {}


def mx(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for mode, ctas in variants:
    os.environ["B200KS_HALO"] = mode
    os.environ["B200KS_PUSH_CTAS"] = ctas
    ids = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = api.Context(dims, device=local_rank, grid=grid, rank=rank, nranks=world, nccl_id=ids[0])
    ctx.links_synthetic(1234)
    vb, vx = ctx.vec_create(), ctx.vec_create()
    ctx.vec_gaussian(vb, EVEN, 5678)
    row = {"halo_mode": ctx.halo_mode()}
    for prec in (2, 1):
        dist.barrier()
        row["dslash_ms_f%d" % (32 * prec)] = mx(ctx.dslash_time(prec, EVEN, 100))
    for mixed in (0, 1):
        for rep in range(2):
            ctx.vec_zero(vx, EVEN)
            dist.barrier()
            it, res = ctx.congrad_dev(vb, vx, 0.05, EVEN, 300, 1, 1e-10, mixed_precision=mixed)
        row["cg_us_per_iter_mixed%d" % mixed] = mx(1e6 * res["device_seconds"] / it)
    out["%s/%s" % (mode, ctas)] = row
    if rank == 0:
        print(mode, ctas, row, file=sys.stderr, flush=True)
    ctx.close()
    dist.barrier()
if rank == 0:
    print(json.dumps({"n": world, "grid": grid, "rows": out}, indent=1))
dist.destroy_process_group()
