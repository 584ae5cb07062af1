"""Host-side data-parallel plumbing (one process per GPU).  Pure torch / torch.distributed on whatever
backend the process group uses (NCCL on the GPU box, gloo in the CPU tests): nothing here touches libsnb200.

  * gradient exchange: one all-reduce (sum) over the LIVE prefix of the flat gradient buffer
    [ MLP block | hash-table levels < n_active ] -- levels that are not active yet have exactly zero gradient on every
    rank, so they are neither reduced nor swept by Adam;
  * patch sharding: rank r draws its own patches from the counter-based stream (seed + 7919 r, step) -- weak scaling,
    no data-path collective;
  * mesh extraction: x-slabs of the SDF lattice per rank (+1 halo plane), gathered and welded on rank 0.

The reference has no multi-GPU code at all; this is new functionality around the same operators.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

RANK_SEED_STRIDE = 7919


def rank_seed(seed: int, rank: int) -> int:
    return seed + RANK_SEED_STRIDE * rank


def live_numel(small_pad: int, level_offsets: Sequence[int], n_active: int) -> int:
    """Floats at the front of the flat parameter/gradient buffers that can be non-zero: MLP block + active table levels
    (2 features per entry)."""
    return small_pad + 2 * int(level_offsets[n_active])


def allreduce_live_gradients(flat_grad: torch.Tensor, n_live: int, world_size: int) -> float:
    """Sum-all-reduce flat_grad[:n_live] in place; returns the scale (1/world) Adam applies to the summed gradient."""
    if world_size <= 1:
        return 1.0
    import torch.distributed as dist
    dist.all_reduce(flat_grad[:n_live])
    return 1.0 / world_size


# ------------------------------------------------------------------------------------------------
# mesh extraction: slab partition + gather/weld
# ------------------------------------------------------------------------------------------------
def slab_cells(resolution: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Cell range [c0, c1) along x owned by `rank`; the rank evaluates lattice planes c0 .. c1 inclusive (c1 is the halo
    plane it shares with rank+1).  resolution lattice points => resolution-1 cells."""
    n_cells = resolution - 1
    base, rem = divmod(n_cells, world_size)
    c0 = rank * base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)


def merge_slab_meshes(parts: List[Tuple[torch.Tensor, torch.Tensor, int]]):
    """parts[r] = (vertices [V_r,3], triangles [T_r,3] (local ids), n_main_r), ordered by rank.  Each slab numbers the
    vertices of its LAST lattice plane after everything else (ids >= n_main_r); for every slab but the last one that
    plane is the next slab's FIRST plane, whose vertices the next slab numbers from 0 in the same (y,z) order -- so halo
    vertices are dropped and their ids rebased onto the neighbour instead of being welded by position."""
    bases, total = [], 0
    for r, (v, t, n_main) in enumerate(parts):
        bases.append(total)
        total += int(v.shape[0]) if r == len(parts) - 1 else int(n_main)
    verts, tris = [], []
    for r, (v, t, n_main) in enumerate(parts):
        last = r == len(parts) - 1
        verts.append(v if last else v[:n_main])
        t = t.to(torch.int64)
        if last:
            tris.append(t + bases[r])
        else:
            tris.append(torch.where(t >= n_main, t - n_main + bases[r + 1], t + bases[r]))
    return torch.cat(verts, 0), torch.cat(tris, 0)


def gather_variable(t: torch.Tensor, dst: int = 0) -> Optional[List[torch.Tensor]]:
    """Gather tensors whose first dimension differs per rank onto `dst` (padded all_gather: works on gloo and NCCL)."""
    import torch.distributed as dist
    world = dist.get_world_size()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    if dist.get_rank() != dst:
        return None
    return [o[:s] for o, s in zip(out, sizes)]


def gather_slab_meshes(vertices: torch.Tensor, triangles: torch.Tensor, n_main: int, dst: int = 0):
    """Collective: every rank passes its slab mesh; rank `dst` gets the merged (vertices, triangles), the others None."""
    import torch.distributed as dist
    vs = gather_variable(vertices, dst)
    ts = gather_variable(triangles, dst)
    nm = gather_variable(torch.tensor([n_main], dtype=torch.int64, device=vertices.device), dst)
    if dist.get_rank() != dst:
        return None
    return merge_slab_meshes([(v, t, int(m.item())) for v, t, m in zip(vs, ts, nm)])
