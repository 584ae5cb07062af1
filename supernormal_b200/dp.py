"""Host-side data-parallel plumbing (one process per GPU).  Pure torch / torch.distributed on whatever
backend the process group uses (NCCL on the GPU box, gloo in the CPU tests): nothing here touches libsnb200.

  * gradient exchange over the LIVE prefix of the flat gradient buffer [ MLP block | hash-table levels < n_active ] --
    levels that are not active yet have exactly zero gradient on every rank, so they are neither reduced nor swept by
    Adam.  Default on a GPU box: the peer-memory step tail (libsnb200's snb_train_tail_peer; PeerGroup below sets up the
    symmetric allocation and the static chunk ownership); fallback / gloo tests: one all-reduce (sum);
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
# peer-memory gradient exchange (NVLink / NVSwitch): symmetric allocation + static chunk ownership
# ------------------------------------------------------------------------------------------------
PEER_CHUNK_FLOATS = 4096     # 1024 float4: one CTA pass of train_tail_peer_kernel
PEER_FLAG_WORDS = 32         # u32 per rank: [0,8) start flags, [8,16) done flags, [16] error code
MAX_PEERS = 8


def chunk_owner(chunk: int, world_size: int) -> int:
    """Rank that reduces / updates / broadcasts table chunk `chunk` (static: Adam state of a parameter never moves, whatever the
    number of active levels)."""
    return chunk % world_size


def owned_floats(n_live_table: int, rank: int, world_size: int) -> int:
    """Floats of the live table range [0, n_live_table) whose Adam state lives on `rank`."""
    n_chunks = -(-n_live_table // PEER_CHUNK_FLOATS)
    total = 0
    for c in range(rank, n_chunks, world_size):
        total += min(PEER_CHUNK_FLOATS, n_live_table - c * PEER_CHUNK_FLOATS)
    return total


def owner_mask(n: int, rank: int, world_size: int, device=None) -> torch.Tensor:
    """bool[n]: table floats whose chunk is owned by `rank` (chunk c = floats [4096 c, 4096 (c+1)))."""
    return (torch.arange(n, device=device) // PEER_CHUNK_FLOATS) % world_size == rank


def carve_layout(sizes_bytes: Sequence[int], align: int = 256) -> Tuple[List[int], int]:
    """Offsets of consecutive sub-buffers inside one allocation (each `align`-byte aligned) and the total size."""
    offs, cur = [], 0
    for nb in sizes_bytes:
        cur = (cur + align - 1) // align * align
        offs.append(cur)
        cur += int(nb)
    return offs, (cur + align - 1) // align * align


class PeerGroup:
    """One symmetric allocation per rank (torch.distributed._symmetric_memory: CUDA VMM memory mapped into every peer of the
    box), carved into the buffers the peer tail kernel needs.  `ptrs(off)` = the device address of byte offset `off` in every
    rank's allocation, as seen from THIS rank.  Raises if symmetric memory cannot be set up (the trainer then keeps NCCL)."""

    def __init__(self, nbytes: int, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = group or dist.group.WORLD
        self.buf = symm.empty(int(nbytes), dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, group)
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        if self.world > MAX_PEERS:
            raise RuntimeError(f"peer tail supports up to {MAX_PEERS} ranks")
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        own = self.buf.data_ptr()
        off = own - ptrs[self.rank]          # the tensor may sit at an offset inside the exchanged allocation
        if not (0 <= off and off + int(nbytes) <= int(self.hdl.buffer_size)):
            raise RuntimeError("symmetric memory: tensor is not inside the exchanged allocation")
        self.base = [p + off for p in ptrs]
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)                  # nobody signals into a peer's flags before they are zeroed

    def view(self, off: int, numel: int, dtype: torch.dtype) -> torch.Tensor:
        nb = numel * torch.empty((), dtype=dtype).element_size()
        return self.buf[off:off + nb].view(dtype)

    def ptrs(self, off: int) -> List[int]:
        return [b + off for b in self.base]


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
