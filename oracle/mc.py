"""numpy restatement of the mesh-extraction tail (models/renderer.py:9-34: extract_fields + mcubes.marching_cubes +
vertex rescale).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED: PyMCubes 0.1.4 (create_env.sh:14) is neither vendored under /root/reference nor installed here, so its
vertex/triangle ORDER cannot be reproduced; what is restated is the published algorithm (Lorensen & Cline 1987: 256-case
table lookup, linear interpolation of the crossing along each cube edge, one shared vertex per crossed lattice edge).
The case tables come from scripts/gen_mc_tables.py.  The numbering below is the contract of supernormal_b200/csrc/mesh.cu:

  vertices:  [ y/z-edge vertices of lattice planes 0..nx-2, lattice order (x-major), y before z ]
             [ x-edge vertices, lattice order ]
             [ y/z-edge vertices of the LAST plane nx-1 ]         <- the halo plane of an x-slab (dropped when merging)
  triangles: cells in lattice order, table order inside a cell; vertex winding such that normals point from
             u > iso (inside: the reference passes u = -sdf) towards u <= iso.
"""
from __future__ import annotations

import numpy as np

from .mc_tables import NUM_TRIS, TRI_TABLE


def marching_cubes(u: np.ndarray, iso: float = 0.0, x_offset: int = 0):
    """u: float32 [nx, ny, nz] -> (vertices f32 [V,3] in lattice-index coordinates (+x_offset), triangles i32 [T,3],
    n_main, n_last)."""
    u = np.ascontiguousarray(u, dtype=np.float32)
    nx, ny, nz = u.shape
    iso = np.float32(iso)
    inside = u > iso
    cx = np.zeros_like(inside)
    cy = np.zeros_like(inside)
    cz = np.zeros_like(inside)
    cx[:-1] = inside[:-1] != inside[1:]
    cy[:, :-1] = inside[:, :-1] != inside[:, 1:]
    cz[:, :, :-1] = inside[:, :, :-1] != inside[:, :, 1:]
    cntA = cy.astype(np.int64) + cz
    plane = ny * nz
    counts = np.concatenate([cntA[:nx - 1].ravel(), cx[:nx - 1].astype(np.int64).ravel(), cntA[nx - 1].ravel()])
    off = np.concatenate([[0], np.cumsum(counts)])
    n_total = int(off[-1])
    n_main = int(off[2 * (nx - 1) * plane])
    offA = np.empty((nx, ny, nz), np.int64)
    offA[:nx - 1] = off[:(nx - 1) * plane].reshape(nx - 1, ny, nz)
    offA[nx - 1] = off[2 * (nx - 1) * plane:2 * (nx - 1) * plane + plane].reshape(ny, nz)
    offB = np.zeros((nx, ny, nz), np.int64)
    offB[:nx - 1] = off[(nx - 1) * plane:2 * (nx - 1) * plane].reshape(nx - 1, ny, nz)

    verts = np.zeros((n_total, 3), np.float32)
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")

    def emit(mask, ids, axis):
        i, j, k = ii[mask], jj[mask], kk[mask]
        v0 = u[i, j, k]
        v1 = u[i + (axis == 0), j + (axis == 1), k + (axis == 2)]
        mu = ((iso - v0) / (v1 - v0)).astype(np.float32)
        p = np.stack([i + x_offset, j, k], -1).astype(np.float32)   # global lattice index first: slab-invariant rounding
        p[:, axis] = p[:, axis] + mu
        verts[ids[mask]] = p
    emit(cy, offA, 1)
    emit(cz, offA + cy, 2)
    emit(cx, offB, 0)

    # cells
    ins = inside.astype(np.int64)
    case = np.zeros((nx - 1, ny - 1, nz - 1), np.int64)
    for c in range(8):
        dx, dy, dz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        case |= ins[dx:nx - 1 + dx, dy:ny - 1 + dy, dz:nz - 1 + dz] << c
    ntri = NUM_TRIS[case]
    ci, cj, ck = np.nonzero(ntri)
    tris = []
    if ci.size:
        cs = case[ci, cj, ck]
        for t in range(int(ntri.max())):
            sel = NUM_TRIS[cs] > t
            i, j, k, c = ci[sel], cj[sel], ck[sel], cs[sel]
            tri = np.empty((i.size, 3), np.int64)
            for a in range(3):
                e = TRI_TABLE[c, 3 * t + a]
                axis, q = e // 4, e % 4
                lo, hi = q & 1, q >> 1
                ox = np.where(axis == 0, 0, lo)
                oy = np.where(axis == 0, lo, np.where(axis == 1, 0, hi))
                oz = np.where(axis == 2, 0, hi)
                pi, pj, pk = i + ox, j + oy, k + oz
                vid = np.where(axis == 0, offB[pi, pj, pk], np.where(axis == 1, offA[pi, pj, pk], offA[pi, pj, pk] + cy[pi, pj, pk]))
                tri[:, a] = vid
            # order key: (cell lattice index, t)
            tris.append((((i * (ny - 1) + j) * (nz - 1) + k) * 8 + t, tri))
        keys = np.concatenate([k for k, _ in tris])
        allt = np.concatenate([t for _, t in tris])
        tris = allt[np.argsort(keys, kind="stable")]
    else:
        tris = np.zeros((0, 3), np.int64)
    return verts, tris.astype(np.int32), n_main, n_total - n_main


def extract_fields(bound_min, bound_max, resolution, query_func):
    """models/renderer.py:9-23 without the 64^3 chunking (which does not change values): u[i,j,k] = query(x_i, y_j, z_k)."""
    import torch
    X = torch.linspace(bound_min[0], bound_max[0], resolution)
    Y = torch.linspace(bound_min[1], bound_max[1], resolution)
    Z = torch.linspace(bound_min[2], bound_max[2], resolution)
    xx, yy, zz = torch.meshgrid(X, Y, Z, indexing="ij")
    pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1)
    return query_func(pts).reshape(resolution, resolution, resolution).numpy().astype(np.float32)


def rescale(vertices, bound_min, bound_max, resolution):
    """models/renderer.py:33."""
    bmin, bmax = np.asarray(bound_min, np.float64), np.asarray(bound_max, np.float64)
    return vertices.astype(np.float64) / (resolution - 1.0) * (bmax - bmin)[None, :] + bmin[None, :]   # mcubes returns float64 vertices


def mesh_checks(vertices, triangles):
    """closedness / orientation statistics of a triangle mesh: (n_boundary_edges, n_nonmanifold_edges, n_misoriented_edges, signed_volume)."""
    t = triangles.astype(np.int64)
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    key_dir = e[:, 0] * (t.max() + 1 if t.size else 1) + e[:, 1]
    und = np.sort(e, 1)
    key = und[:, 0] * (t.max() + 1 if t.size else 1) + und[:, 1]
    _, cnt = np.unique(key, return_counts=True)
    _, cnt_dir = np.unique(key_dir, return_counts=True)
    p = vertices.astype(np.float64)
    vol = float(np.einsum("ij,ij->i", p[t[:, 0]], np.cross(p[t[:, 1]], p[t[:, 2]])).sum() / 6.0) if t.size else 0.0
    return int((cnt == 1).sum()), int((cnt > 2).sum()), int((cnt_dir > 1).sum()), vol
