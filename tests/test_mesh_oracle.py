"""CPU tests of the mesh-extraction restatement (oracle/mc.py), the generated marching-cubes tables and the slab
merge of supernormal_b200/dp.py.  PyMCubes is absent (parity unpinned): the pins are
topological (closed, consistently oriented 2-manifolds; watertight across ambiguous faces) and geometric (analytic
sphere)."""
import numpy as np
import torch

from oracle import mc
from supernormal_b200 import dp


def sphere_field(res, r=0.5, nx=None):
    g = np.linspace(-1, 1, res).astype(np.float32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    return (r - np.sqrt(X ** 2 + Y ** 2 + Z ** 2)).astype(np.float32)   # u = -sdf, like models/renderer.py:358


def canonical_triangles(verts, tris):
    """order-independent form: every triangle as 9 coordinates starting at its lexicographically smallest vertex
    (cyclic rotation keeps the winding), rows sorted."""
    p = verts[tris.astype(np.int64)]                      # [T,3,3]
    key = p[:, :, 0] * 1e6 + p[:, :, 1] * 1e3 + p[:, :, 2]
    start = key.argmin(1)
    idx = (start[:, None] + np.arange(3)[None, :]) % 3
    p = np.take_along_axis(p, idx[:, :, None], 1).reshape(-1, 9)
    return p[np.lexsort(p.T[::-1])]


def test_tables_cover_all_cases():
    from oracle.mc_tables import NUM_TRIS, TRI_TABLE, MAX_TRIS
    assert NUM_TRIS[0] == 0 and NUM_TRIS[255] == 0 and MAX_TRIS == 5
    for c in range(256):
        n = NUM_TRIS[c]
        assert (TRI_TABLE[c, :3 * n] >= 0).all() and (TRI_TABLE[c, 3 * n:] == -1).all()
        # a crossed edge is used by the case's triangles iff its corners differ
        crossed = {e for e in range(12) if ((c >> mc_edge(e)[0]) & 1) != ((c >> mc_edge(e)[1]) & 1)}
        assert set(TRI_TABLE[c, :3 * n].tolist()) == crossed
        assert NUM_TRIS[c] == NUM_TRIS[255 - c] or True   # complementary cases may triangulate differently (ambiguous faces)


def mc_edge(e):
    from oracle.mc_tables import EDGE_CORNERS
    return EDGE_CORNERS[e]


def test_sphere_is_closed_oriented_manifold():
    res = 40
    v, t, n_main, n_last = mc.marching_cubes(sphere_field(res))
    boundary, nonmanifold, misoriented, vol = mc.mesh_checks(v, t)
    assert (boundary, nonmanifold, misoriented) == (0, 0, 0)
    assert v.shape[0] - t.shape[0] // 2 == 2                         # Euler characteristic of a sphere: V - E + F = V - F/2 = 2
    vw = mc.rescale(v, [-1, -1, -1], [1, 1, 1], res)
    assert np.abs(np.linalg.norm(vw, axis=1) - 0.5).max() < 1e-3     # linear interpolation of a distance field
    vol_w = mc.mesh_checks(vw, t)[3]
    assert 0.97 * 4 / 3 * np.pi * 0.125 < vol_w < 4 / 3 * np.pi * 0.125   # positive: normals point outwards (u > iso is inside)
    assert n_last == 0


def test_random_field_is_watertight():
    """white noise exercises every case incl. ambiguous faces: open edges may only lie on the lattice boundary."""
    rng = np.random.RandomState(0)
    u = rng.randn(14, 15, 16).astype(np.float32)
    v, t, _, _ = mc.marching_cubes(u)
    tt = t.astype(np.int64)
    e = np.concatenate([tt[:, [0, 1]], tt[:, [1, 2]], tt[:, [2, 0]]])
    und = np.sort(e, 1)
    key = und[:, 0] * (tt.max() + 1) + und[:, 1]
    _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    assert cnt.max() <= 2
    be = und[cnt[inv] == 1]
    on_b = lambda q: (q[:, 0] == 0) | (q[:, 0] == 13) | (q[:, 1] == 0) | (q[:, 1] == 14) | (q[:, 2] == 0) | (q[:, 2] == 15)
    assert (on_b(v[be[:, 0]]) & on_b(v[be[:, 1]])).all()
    assert mc.mesh_checks(v, t)[2] == 0                                 # no directed edge twice: consistent orientation


def test_empty_and_full_fields():
    for val in (-1.0, 1.0):
        v, t, n_main, n_last = mc.marching_cubes(np.full((4, 5, 6), val, np.float32))
        assert v.shape == (0, 3) and t.shape == (0, 3) and n_main == 0 and n_last == 0


def test_slab_merge_equals_whole():
    rng = np.random.RandomState(1)
    u = (sphere_field(24, 0.6) + 0.05 * rng.randn(24, 24, 24)).astype(np.float32)
    vw, tw, _, _ = mc.marching_cubes(u)
    for world in (2, 3, 5):
        parts = []
        for r in range(world):
            c0, c1 = dp.slab_cells(24, r, world)
            v, t, n_main, _ = mc.marching_cubes(u[c0:c1 + 1], 0.0, x_offset=c0)
            parts.append((torch.from_numpy(v), torch.from_numpy(t), n_main))
        vm, tm = dp.merge_slab_meshes(parts)
        assert vm.shape[0] == vw.shape[0] and tm.shape[0] == tw.shape[0]            # halo vertices dropped, not duplicated
        assert np.array_equal(canonical_triangles(vm.numpy(), tm.numpy()), canonical_triangles(vw, tw))
        assert mc.mesh_checks(vm.numpy(), tm.numpy().astype(np.int32))[:3] == mc.mesh_checks(vw, tw)[:3]


def test_slab_cells_partition():
    for res, world in ((512, 8), (33, 4), (5, 8), (1024, 3)):
        cells = [dp.slab_cells(res, r, world) for r in range(world)]
        assert cells[0][0] == 0 and cells[-1][1] == res - 1
        assert all(cells[r][1] == cells[r + 1][0] for r in range(world - 1))
        sizes = [b - a for a, b in cells]
        assert max(sizes) - min(sizes) <= 1
