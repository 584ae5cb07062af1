"""CPU tests of supernormal_b200/mesh_post.py: largest edge-connected cluster
(exp_runner.py:508-524) against a union-find oracle, SDF sphere tracing / visible points (exp_runner.py:580-592)
against an analytic sphere.  open3d / trimesh are not in the image: parity unpinned, the oracle is the definition."""
import numpy as np
import pytest
import torch

from supernormal_b200 import mesh_post


def _uf_clusters(tri: np.ndarray) -> np.ndarray:
    """union-find over triangles that share an (undirected) edge -> root id per triangle"""
    parent = list(range(len(tri)))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    first = {}
    for t, (a, b, c) in enumerate(tri):
        for e in ((a, b), (b, c), (c, a)):
            k = (min(e), max(e))
            if k in first:
                ra, rb = find(first[k]), find(t)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
            else:
                first[k] = t
    return np.array([find(i) for i in range(len(tri))])


def _grid_mesh(nx, ny, v0):
    """triangulated nx x ny quad grid, vertex ids starting at v0"""
    idx = np.arange((nx + 1) * (ny + 1)).reshape(ny + 1, nx + 1) + v0
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    return np.concatenate([np.stack([a, b, c], 1), np.stack([b, d, c], 1)]), (nx + 1) * (ny + 1)


def _test_mesh(seed):
    rng = np.random.RandomState(seed)
    parts, v0 = [], 0
    for nx, ny in ((30, 20), (7, 5), (3, 3), (1, 1)):
        t, nv = _grid_mesh(nx, ny, v0)
        parts.append(t)
        v0 += nv
    # two triangles touching the big grid in ONE vertex only (vertex-connected, not edge-connected): separate clusters
    parts.append(np.array([[0, v0, v0 + 1], [5, v0 + 2, v0 + 3]]))
    v0 += 4
    tri = np.concatenate(parts)
    tri = tri[rng.permutation(len(tri))]
    return rng.randn(v0 + 3, 3).astype(np.float32), tri.astype(np.int32)   # + 3 unreferenced vertices


@pytest.mark.parametrize("seed", [0, 1])
def test_triangle_clusters_match_union_find(seed):
    v, tri = _test_mesh(seed)
    lab, counts = mesh_post.triangle_clusters(torch.from_numpy(tri))
    ref = _uf_clusters(tri)
    lab = lab.numpy()
    # same partition: labels are in one-to-one correspondence
    pairs = set(zip(lab.tolist(), ref.tolist()))
    assert len(pairs) == len(set(lab.tolist())) == len(set(ref.tolist())) == 6
    assert sorted(counts.tolist()) == sorted(np.unique(ref, return_counts=True)[1].tolist()) == [1, 1, 2, 18, 70, 1200]


def test_remove_isolated_clusters_keeps_largest_and_reindexes():
    v, tri = _test_mesh(2)
    v2, t2 = mesh_post.remove_isolated_clusters(v, tri)
    assert isinstance(v2, np.ndarray) and t2.dtype == tri.dtype
    assert t2.shape == (1200, 3) and v2.shape == (31 * 21, 3)
    # geometry of the kept triangles is unchanged, vertex order preserved, every vertex referenced
    ref = _uf_clusters(tri)
    big = np.bincount(ref).argmax()
    assert np.array_equal(v2[t2], v[tri[ref == big]])
    assert np.array_equal(np.unique(t2), np.arange(len(v2)))
    assert np.array_equal(v2, v[:31 * 21])
    # torch in -> torch out; empty mesh passes through
    v3, t3 = mesh_post.remove_isolated_clusters(torch.from_numpy(v), torch.from_numpy(tri))
    assert torch.is_tensor(v3) and np.array_equal(t3.numpy(), t2) and np.array_equal(v3.numpy(), v2)
    ve, te = mesh_post.remove_isolated_clusters(v, np.zeros((0, 3), np.int32))
    assert te.shape == (0, 3) and ve.shape == v.shape


def test_sphere_trace_and_visible_points_on_analytic_sphere():
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene
    ds = SyntheticDataset(SyntheticScene(n_views=3, H=48, W=64, exclude_views=()), device="cpu", with_v_inverse=False)
    r = float(ds.scene.radius)
    sdf = lambda x: x.norm(dim=-1, keepdim=True) - r
    pts = mesh_post.find_visible_points(ds, sdf)
    n_fg = int((ds.masks > 0.5).sum())
    assert pts.shape[1] == 3 and 0.97 * n_fg <= pts.shape[0] <= n_fg      # grazing rays at the silhouette may miss
    assert (pts.norm(dim=-1) - r).abs().max() < 2e-4
    # the point is the FIRST hit: on the camera side of the sphere
    o, d = mesh_post.view_rays_within_mask(ds, 1)
    near, far = ds.near_far_from_sphere(o, d)
    x, hit = mesh_post.sphere_trace(sdf, o, d, near, far)
    assert hit.float().mean() > 0.97
    assert ((x[hit] - o[hit]) * x[hit]).sum(-1).max() < 0        # outward normal faces the camera
    # rays that miss the unit sphere (NaN interval) are misses, not NaN points
    d_miss = torch.nn.functional.normalize(torch.tensor([[0.0, 0.0, 1.0]]), dim=-1)
    o_miss = torch.tensor([[5.0, 0.0, 0.0]])
    n2, f2 = ds.near_far_from_sphere(o_miss, d_miss)
    x2, h2 = mesh_post.sphere_trace(sdf, o_miss, d_miss, n2, f2)
    assert not h2.any() and torch.isfinite(x2).all()
