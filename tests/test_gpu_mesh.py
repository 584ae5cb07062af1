"""GPU parity of the mesh-extraction path: lattice SDF query and device marching cubes through the
C ABI against oracle/mc.py (exact numbering contract -> array equality), plus the reference-shaped
extract_geometry on a model (closed sphere-like mesh, slab-invariance)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import mc

pytestmark = pytest.mark.gpu


def _field(kind, shape, seed=0):
    rng = np.random.RandomState(seed)
    if kind == "noise":
        return rng.randn(*shape).astype(np.float32)
    g = [np.linspace(-1, 1, n).astype(np.float32) for n in shape]
    X, Y, Z = np.meshgrid(*g, indexing="ij")
    u = (0.55 - np.sqrt(X ** 2 + Y ** 2 + Z ** 2)).astype(np.float32)
    if kind == "bumpy":
        u += (0.04 * rng.randn(*shape)).astype(np.float32)
    return u


@pytest.mark.parametrize("kind,shape,iso,xoff", [("sphere", (40, 40, 40), 0.0, 0), ("noise", (17, 19, 23), 0.1, 0), ("bumpy", (9, 64, 48), 0.0, 37),
                                                 ("noise", (2, 2, 2), 0.0, 0), ("sphere", (130, 70, 66), 0.02, 5)])
def test_marching_cubes_matches_oracle(cuda, kind, shape, iso, xoff):
    from supernormal_b200 import mesh
    u = _field(kind, shape)
    vo, to, n_main_o, _ = mc.marching_cubes(u, iso, x_offset=xoff)
    v, t, n_main = mesh.marching_cubes(torch.from_numpy(u).to(cuda), iso, x_offset=xoff)
    assert n_main == n_main_o and v.shape[0] == vo.shape[0] and t.shape[0] == to.shape[0]
    assert np.array_equal(t.cpu().numpy(), to)                       # same numbering contract: identical triangles
    assert np.array_equal(v.cpu().numpy(), vo)                       # fp32 (iso - v0) / (v1 - v0), IEEE ops on both sides


def test_marching_cubes_empty_and_errors(cuda):
    from supernormal_b200 import mesh
    v, t, n_main = mesh.marching_cubes(torch.full((5, 6, 7), -1.0, device=cuda))
    assert v.shape == (0, 3) and t.shape == (0, 3) and n_main == 0
    with pytest.raises(ValueError):
        mesh.marching_cubes(torch.zeros(1, 4, 4, device=cuda))
    with pytest.raises(NotImplementedError):
        mesh.marching_cubes(torch.zeros(4, 4, 4))


def _model(cuda, n_active=4):
    from supernormal_b200.trainer import SDFModel
    from supernormal_b200.synthetic import DILIGENT_CONF
    m = SDFModel(DILIGENT_CONF["encoding"], 0.6, 0.5, device=cuda)
    with torch.no_grad():   # make the hash grid matter
        m.table.copy_((torch.rand(m.n_table, device=cuda) * 2 - 1) * 0.02)
    m.refresh_table_f16()
    m.n_active = n_active
    m.prep()
    return m


def test_grid_query_equals_point_query(cuda):
    """extract_fields (models/renderer.py:9-23): same values as querying the meshgrid points one by one."""
    from supernormal_b200 import mesh
    m = _model(cuda)
    res = 37
    u = mesh.extract_fields(m, [-1, -1, -1], [1, 1, 1], res)
    X = torch.linspace(-1, 1, res, device=cuda)
    xx, yy, zz = torch.meshgrid(X, X, X, indexing="ij")
    pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1)
    ref = -m.sdf(pts).view(res, res, res)
    # not bit-equal: softplus takes its saturation shortcut per warp, and the two kernels group points into warps differently (4e-10 per hidden unit -> a few ulp of the sum)
    assert torch.allclose(u, ref, rtol=0, atol=1e-6)
    sl = mesh.extract_fields(m, [-1, -1, -1], [1, 1, 1], res, x_range=(11, 20))
    assert torch.allclose(sl, ref[11:20], rtol=0, atol=1e-6)


def test_extract_geometry_sphere_and_slabs(cuda):
    from supernormal_b200 import mesh
    m = _model(cuda, n_active=0)   # geometric init: sdf ~ |x| - 0.6
    res = 96
    B0, B1 = [-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]   # the geometric-init "sphere" (radius ~0.6, +-20 %) must not touch the lattice boundary
    v, t = mesh.extract_geometry(m, B0, B1, res, 0.0, distributed=False)
    assert v.dtype == np.float64 and t.dtype == np.int32
    boundary, nonmanifold, misoriented, vol = mc.mesh_checks(v, t)
    assert (boundary, nonmanifold, misoriented) == (0, 0, 0) and vol > 0
    r = np.linalg.norm(v, axis=1)
    assert 0.4 < r.mean() < 1.0 and r.std() < 0.25   # geometric init is only roughly a sphere of radius bias = 0.6
    v4, t4 = mesh.extract_geometry(m, B0, B1, res, 0.0, distributed=False, slabs=4)
    assert v4.shape == v.shape and t4.shape == t.shape
    from test_mesh_oracle import canonical_triangles
    assert np.array_equal(canonical_triangles(v4, t4), canonical_triangles(v, t))
    # against the oracle fed with the same field
    u = mesh.extract_fields(m, B0, B1, res).cpu().numpy()
    vo, to, _, _ = mc.marching_cubes(u, 0.0)
    assert np.array_equal(to, t) and np.allclose(mc.rescale(vo, B0, B1, res), v, atol=1e-12)


def test_mesh_512_properties(cuda):
    """BASELINE config 5 size: 512^3 lattice in 8 sequential x-slabs; size-independent properties."""
    from supernormal_b200 import mesh
    m = _model(cuda, n_active=0)
    v, t = mesh.extract_geometry(m, [-1.5, -1.5, -1.5], [1.5, 1.5, 1.5], 512, 0.0, distributed=False, slabs=8)
    boundary, nonmanifold, misoriented, vol = mc.mesh_checks(v, t)
    assert (boundary, nonmanifold, misoriented) == (0, 0, 0)
    assert v.shape[0] - t.shape[0] // 2 == 2      # one closed genus-0 surface
    r = np.linalg.norm(v, axis=1)
    assert 0.5 * 4 / 3 * np.pi * r.mean() ** 3 < vol < 1.5 * 4 / 3 * np.pi * r.mean() ** 3
