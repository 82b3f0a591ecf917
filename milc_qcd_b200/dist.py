"""Host-side decomposition logic for one-rank-per-GPU runs.

The lattice is split in t first, then z (BASELINE.json north_star), each rank owning a
hypercubic sub-lattice stored in MILC's own per-node order -- what a MILC MPI rank holds under
``generic/layout_hyper_prime.c:186-229,509-520`` (local lexicographic index, even sites first).
Everything here is numpy index arithmetic; the exchange itself is done by the CUDA library
(NCCL send/recv of depth-3 ghost zones, all-reduce of the CG scalars).
"""
import numpy as np

from .fields import lex_to_milc


def rank_grid(nranks):
    """{1,1,gz,gt}: t is split first (up to 4 ways), then z."""
    gt = 1
    while gt < 4 and nranks % (gt * 2) == 0:
        gt *= 2
    gz = nranks // gt
    if gz * gt != nranks:
        raise ValueError("unsupported rank count %d" % nranks)
    return (1, 1, gz, gt)


def rank_coords(grid, rank):
    """Grid coordinates of a rank (t slowest), matching b200ks_create_dist."""
    return (0, 0, rank % grid[2], rank // grid[2])


def local_dims(dims, grid):
    for d in range(4):
        if dims[d] % grid[d]:
            raise ValueError("lattice extent %d not divisible by grid %d" % (dims[d], grid[d]))
    return tuple(dims[d] // grid[d] for d in range(4))


def local_to_global_index(dims, grid, rank):
    """idx[i_local_milc] = i_global_milc for the sub-lattice of `rank`."""
    L = local_dims(dims, grid)
    co = rank_coords(grid, rank)
    org = [co[d] * L[d] for d in range(4)]
    Vl = int(np.prod(L))
    lex = np.arange(Vl, dtype=np.int64)
    x = lex % L[0] + org[0]
    y = (lex // L[0]) % L[1] + org[1]
    z = (lex // (L[0] * L[1])) % L[2] + org[2]
    t = lex // (L[0] * L[1] * L[2]) + org[3]
    glex = x + dims[0] * (y + dims[1] * (z + dims[2] * t))
    gperm = lex_to_milc(dims)          # global lex -> global milc
    lperm = lex_to_milc(L)             # local lex  -> local milc
    out = np.empty(Vl, dtype=np.int64)
    out[lperm] = gperm[glex]
    return out


def scatter_field(field, dims, grid, rank):
    """Local sub-lattice (local MILC order) of a global MILC-order field (first axis = sites)."""
    return np.ascontiguousarray(field[local_to_global_index(dims, grid, rank)])


def gather_field(parts, dims, grid):
    """Inverse of scatter_field: assemble the global field from all ranks' local fields."""
    n = len(parts)
    V = int(np.prod(dims))
    out = np.empty((V,) + parts[0].shape[1:], dtype=parts[0].dtype)
    for r in range(n):
        out[local_to_global_index(dims, grid, r)] = parts[r]
    return out


def ghost_sites(dims, grid, rank, d, side, parity_bit):
    """Global MILC indices of the depth-3 ghost zone of `rank` in direction d (2=z, 3=t),
    side 0 = behind (local coords -3..-1), 1 = ahead (L..L+2), for sites of the given parity,
    in the order the device ghost buffer uses: [slice][t][z or nothing][y][xh]."""
    L = local_dims(dims, grid)
    co = rank_coords(grid, rank)
    org = [co[k] * L[k] for k in range(4)]
    gperm = lex_to_milc(dims)
    out = []
    for s in range(3):
        cd = (-3 + s) if side == 0 else (L[d] + s)
        rng = [range(L[0]), range(L[1]), range(L[2]), range(L[3])]
        rng[d] = [cd]
        for t in rng[3]:
            for z in rng[2]:
                for y in rng[1]:
                    for x in rng[0]:
                        if (x + y + z + t) & 1 != parity_bit:
                            continue
                        g = [(x + org[0]) % dims[0], (y + org[1]) % dims[1], (z + org[2]) % dims[2], (t + org[3]) % dims[3]]
                        out.append(gperm[g[0] + dims[0] * (g[1] + dims[1] * (g[2] + dims[2] * g[3]))])
    return np.array(out, dtype=np.int64)
