"""Mesh extraction on the device: the tail of the reference's time-to-mesh (models/renderer.py:9-34,353-358).

extract_fields / extract_geometry keep the reference's names, argument meaning and return convention
((vertices, triangles) as numpy arrays, vertices already mapped to the bounding box, models/renderer.py:26-34), but the
SDF lattice never leaves the GPU: one fused encode+MLP launch per x-slab, then marching cubes in libsnb200
(snb_mc_count / snb_mc_emit).  With torch.distributed initialised the lattice is sharded by x-slab over the ranks and
rank 0 receives the welded mesh (dp.gather_slab_meshes); the other ranks return None.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib, dp
from ._lib import call, ptr


def _axes(bound_min, bound_max, resolution: int, device):
    """the three torch.linspace axes of extract_fields (models/renderer.py:11-13), built with the same op"""
    bmin = [float(v) for v in bound_min]
    bmax = [float(v) for v in bound_max]
    return [torch.linspace(bmin[a], bmax[a], resolution, device=device, dtype=torch.float32) for a in range(3)]


@torch.no_grad()
def extract_fields(model, bound_min, bound_max, resolution: int, x_range: Optional[Tuple[int, int]] = None, mode: int = 2) -> torch.Tensor:
    """u[i,j,k] = -sdf(X[i], Y[j], Z[k]) for lattice planes x_range = [x0, x1) (default: all), as a CUDA tensor
    [x1-x0, res, res].  model: trainer.SDFModel (prep() must reflect the current parameters)."""
    X, Y, Z = _axes(bound_min, bound_max, resolution, model.device)
    x0, x1 = x_range if x_range is not None else (0, resolution)
    xs = X[x0:x1].contiguous()
    out = torch.empty(x1 - x0, resolution, resolution, device=model.device, dtype=torch.float32)
    net = model.net_struct()
    call("snb_sdf_grid_query", ptr(xs), xs.numel(), ptr(Y), resolution, ptr(Z), resolution, C.byref(net), mode, ptr(out))
    return out


@torch.no_grad()
def marching_cubes(u: torch.Tensor, threshold: float = 0.0, x_offset: int = 0):
    """u: CUDA f32 [nx, ny, nz] -> (vertices f32 [V,3] lattice-index coordinates, triangles i32 [T,3], n_main).
    One host read of three counters sizes the outputs (the reference reads the whole field back instead)."""
    if not u.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    u = u.contiguous().float()
    nx, ny, nz = u.shape
    nbytes = _lib.lib().snb_mc_workspace_bytes(nx, ny, nz)
    if nbytes == 0:
        raise ValueError("marching_cubes needs at least 2 lattice points per axis")
    ws = torch.empty(nbytes // 8 + 32, dtype=torch.int64, device=u.device)   # cudaMalloc: 256-byte aligned
    call("snb_mc_count", ptr(u), nx, ny, nz, float(threshold), ptr(ws))
    n_vert, n_main, n_tri = [int(v) for v in ws[:3].tolist()]
    verts = torch.empty(n_vert, 3, dtype=torch.float32, device=u.device)
    tris = torch.empty(n_tri, 3, dtype=torch.int32, device=u.device)
    call("snb_mc_emit", ptr(u), nx, ny, nz, float(threshold), float(x_offset), ptr(ws), n_vert, n_tri,
         ptr(verts) if n_vert else None, ptr(tris) if n_tri else None)
    return verts, tris, n_main


@torch.no_grad()
def extract_geometry(model, bound_min, bound_max, resolution: int, threshold: float = 0.0, distributed: Optional[bool] = None,
                     slabs: int = 1):
    """(vertices float64 [V,3] in bounding-box coordinates, triangles int32 [T,3]) like models/renderer.py:26-34.

    distributed (default: torch.distributed is initialised): each rank extracts its x-slab, rank 0 returns the merged
    mesh and the others None.  slabs > 1 additionally splits a rank's range into sequential slabs (bounds workspace
    memory at 1024^3)."""
    import torch.distributed as dist
    if distributed is None:
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank, world = (dist.get_rank(), dist.get_world_size()) if distributed else (0, 1)
    model.prep()
    c0, c1 = dp.slab_cells(resolution, rank, world)
    parts = []
    # the marching-cubes kernels number potential edge vertices with int32 ids (3 per lattice point of a slab): split this rank's range into
    # enough sequential slabs on its own (1024^3 -- own_objects.conf's val_mesh_res -- needs 2 on one GPU)
    need = -(-(3 * (c1 - c0 + 1) * resolution * resolution) // (2 ** 31 - 1))
    n_sub = max(1, min(max(slabs, need), c1 - c0))
    for s in range(n_sub):
        a = c0 + (c1 - c0) * s // n_sub
        b = c0 + (c1 - c0) * (s + 1) // n_sub
        if b <= a:
            continue
        u = extract_fields(model, bound_min, bound_max, resolution, (a, b + 1))   # + halo plane
        parts.append(marching_cubes(u, threshold, x_offset=a))
        del u
    if parts:
        v, t = dp.merge_slab_meshes(parts)
        n_main = v.shape[0] - (parts[-1][0].shape[0] - parts[-1][2])
    else:   # more ranks than cells
        v, t, n_main = torch.zeros(0, 3, device=model.device), torch.zeros(0, 3, dtype=torch.int64, device=model.device), 0
    if distributed:
        merged = dp.gather_slab_meshes(v, t.to(torch.int64), n_main)
        if merged is None:
            return None
        v, t = merged
    bmin = np.asarray([float(x) for x in bound_min], np.float64)
    bmax = np.asarray([float(x) for x in bound_max], np.float64)
    vertices = v.cpu().numpy().astype(np.float64) / (resolution - 1.0) * (bmax - bmin)[None, :] + bmin[None, :]
    return vertices, t.cpu().numpy().astype(np.int32)
