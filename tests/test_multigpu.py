"""N > 1 coverage.

* CPU (gloo, world_size 2): the host-side decomposition logic -- rank grid, local sub-lattice
  extraction in MILC per-node order, ghost-zone site order -- exercised with a real two-process
  exchange: each rank sends the faces its neighbour needs, and the assembled (local + ghost)
  data must reproduce the oracle's dslash on the global lattice for the rank's own sites.
* GPU (-m gpu, needs >= 2 devices): tests/mgpu_check.py under torch.distributed.run.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

EVEN, ODD, EVENANDODD = 2, 1, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_rank_grid_and_scatter_gather_roundtrip():
    from milc_qcd_b200 import dist as D, fields as F
    assert D.rank_grid(1) == (1, 1, 1, 1)
    assert D.rank_grid(2) == (1, 1, 1, 2)
    assert D.rank_grid(4) == (1, 1, 1, 4)
    assert D.rank_grid(8) == (1, 1, 2, 4)
    dims = (4, 6, 12, 24)  # local extents stay even in every decomposition
    src = F.make_source(dims, seed=3, parity=EVENANDODD)
    for n in (2, 4, 8):
        grid = D.rank_grid(n)
        parts = [D.scatter_field(src, dims, grid, r) for r in range(n)]
        assert all(p.shape[0] == src.shape[0] // n for p in parts)
        assert np.array_equal(D.gather_field(parts, dims, grid), src)
        # local order is MILC's per-node order: even sites first
        L = D.local_dims(dims, grid)
        idx = D.local_to_global_index(dims, grid, n - 1)
        Vl = int(np.prod(L))
        V = int(np.prod(dims))
        assert np.all(idx[:Vl // 2] < V // 2) and np.all(idx[Vl // 2:] >= V // 2)


_GLOO_WORKER = r"""
import os, sys, numpy as np
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
from milc_qcd_b200 import dist as D, fields as F
from oracle.pyoracle import Oracle
EVEN, ODD, EVENANDODD = 2, 1, 3
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
dims = (4, 4, 6, 12)
grid = D.rank_grid(world)
L = D.local_dims(dims, grid)
fat, lng = F.make_links(dims, seed=5)
src = F.make_source(dims, seed=6, parity=EVENANDODD)
lsrc = D.scatter_field(src, dims, grid, rank)
Vl = lsrc.shape[0]
# halo protocol on the host: my high 3 t-slices go to the forward neighbour (its "behind"
# ghost), my low 3 slices to the backward neighbour (its "ahead" ghost); with 2 ranks both
# neighbours are the same peer, messages are matched in issue order like the NCCL group.
peer = 1 - rank
S3h = Vl // 2 // L[3]          # sites per parity per t-slice
ghost = {}
for pbit in (0, 1):
    half = lsrc[pbit * (Vl // 2):(pbit + 1) * (Vl // 2)]
    hi = torch.from_numpy(np.ascontiguousarray(half[(L[3] - 3) * S3h:]))
    lo = torch.from_numpy(np.ascontiguousarray(half[:3 * S3h]))
    behind, ahead = torch.empty_like(hi), torch.empty_like(lo)
    reqs = [dist.isend(hi, peer, tag=10 + pbit), dist.isend(lo, peer, tag=20 + pbit),
            dist.irecv(behind, peer, tag=10 + pbit), dist.irecv(ahead, peer, tag=20 + pbit)]
    for r in reqs: r.wait()
    ghost[(pbit, 0)], ghost[(pbit, 1)] = behind.numpy(), ahead.numpy()
# the received faces must be exactly the global sites the device ghost buffer expects, in its order
for pbit in (0, 1):
    for side in (0, 1):
        want = src[D.ghost_sites(dims, grid, rank, 3, side, pbit)]
        assert np.array_equal(ghost[(pbit, side)], want), (rank, pbit, side)
# and local + ghost data reproduces the oracle's global dslash on this rank's sites:
# build a padded local lattice (t extent + 6) and apply the oracle with the padded links
o = Oracle()
want = D.scatter_field(o.dslash(dims, fat, lng, src, EVENANDODD), dims, grid, rank)
pad_dims = (L[0], L[1], L[2], L[3] + 6)
co = D.rank_coords(grid, rank)
org_t = co[3] * L[3]
gperm = F.lex_to_milc(dims)
pperm = F.lex_to_milc(pad_dims)
Vp = int(np.prod(pad_dims))
lex = np.arange(Vp)
x = lex %% pad_dims[0]; y = (lex // pad_dims[0]) %% pad_dims[1]; z = (lex // (pad_dims[0]*pad_dims[1])) %% pad_dims[2]
t = lex // (pad_dims[0]*pad_dims[1]*pad_dims[2])
gt = (t - 3 + org_t) %% dims[3]
gidx = gperm[x + dims[0]*(y + dims[1]*(z + dims[2]*gt))]
psrc = np.empty((Vp,) + src.shape[1:]); psrc[pperm] = src[gidx]
pfat = np.empty((Vp,) + fat.shape[1:]); pfat[pperm] = fat[gidx]
plng = np.empty((Vp,) + lng.shape[1:]); plng[pperm] = lng[gidx]
# (the ghost slices of psrc are, by the assertion above, exactly what came over the wire)
got_pad = o.dslash(pad_dims, pfat, plng, psrc, EVENANDODD)
inner = (t >= 3) & (t < L[3] + 3)
lidx_lex = (x + L[0]*(y + L[1]*(z + L[2]*(t - 3))))[inner]
lperm = F.lex_to_milc(L)
got = np.empty_like(want); got[lperm[lidx_lex]] = got_pad[pperm[lex[inner]]]
assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
print('GLOO-OK', rank)
dist.destroy_process_group()
"""


def test_halo_protocol_with_gloo_world_size_2(tmp_path):
    script = tmp_path / "gloo_worker.py"
    script.write_text(_GLOO_WORKER % dict(root=ROOT))
    port = _free_port()
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600)
    assert p.stdout.count("GLOO-OK") == 2, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("nproc,dims", [(2, (8, 8, 12, 24)), (4, (8, 8, 8, 24)), (8, (8, 8, 12, 24))])
def test_multigpu_matches_oracle(nproc, dims):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    port = _free_port()
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "mgpu_check.py"), "--dims"] + [str(d) for d in dims],
                       capture_output=True, text=True, timeout=1200)
    assert "MGPU-OK" in p.stdout, p.stdout[-4000:] + p.stderr[-4000:]
