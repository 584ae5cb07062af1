"""GPU parity: nerfacc-shaped operators of supernormal_b200 (through the C ABI) vs the oracle.
Bar: bit-exact for marching / weights (fp32 op order is pinned), exact for packing, 1e-6 relative
for the atomic accumulations (summation order differs, tolerance stated per test)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import torch_ops as T
from conftest import make_march_case

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("kind,step,res,seed", [("shell", 0.01, (128, 128, 128), 0), ("shell", 0.001, (128, 128, 128), 1),
                                                ("random", 0.005, (64, 32, 96), 2), ("full", 0.01, (128, 128, 128), 3),
                                                ("empty", 0.01, (16, 16, 16), 4)])
def test_march_bit_exact(cuda, kind, step, res, seed):
    from supernormal_b200 import nerfacc_api as na
    c = make_march_case(seed=seed, n_rays=1031, res=res, step=step, grid_kind=kind)
    packed_o, ridx_o, t0_o, t1_o = oracle.ray_marching(c["rays_o"], c["rays_d"], c["t_min"], c["t_max"], c["roi"], c["grid"], c["step"])
    packed, ridx, t0, t1 = na._march(_t(c["rays_o"], cuda), _t(c["rays_d"], cuda), _t(c["t_min"], cuda), _t(c["t_max"], cuda),
                                     _t(c["roi"], cuda), _t(c["grid"], cuda), float(c["step"]), 0.0)
    assert np.array_equal(packed.cpu().numpy(), packed_o)
    assert np.array_equal(ridx.cpu().numpy(), ridx_o)
    assert np.array_equal(t0.cpu().numpy().view(np.uint32), t0_o.view(np.uint32))
    assert np.array_equal(t1.cpu().numpy().view(np.uint32), t1_o.view(np.uint32))


def test_march_cone_angle_and_edge_cases(cuda):
    from supernormal_b200 import nerfacc_api as na
    c = make_march_case(seed=7, n_rays=300, step=0.004, grid_kind="random", res=(32, 32, 32))
    for cone in (0.0, 0.004):
        po, ro, a0, a1 = oracle.ray_marching(c["rays_o"], c["rays_d"], c["t_min"], c["t_max"], c["roi"], c["grid"], c["step"], cone)
        p, r, t0, t1 = na._march(_t(c["rays_o"], cuda), _t(c["rays_d"], cuda), _t(c["t_min"], cuda), _t(c["t_max"], cuda),
                                 _t(c["roi"], cuda), _t(c["grid"], cuda), float(c["step"]), cone)
        assert np.array_equal(p.cpu().numpy(), po) and np.array_equal(t0.cpu().numpy().view(np.uint32), a0.view(np.uint32))
        assert np.array_equal(t1.cpu().numpy().view(np.uint32), a1.view(np.uint32))
    # zero rays
    z = torch.zeros((0, 3), device=cuda)
    p, r, t0, t1 = na._march(z, z, z[:, 0], z[:, 0], _t(c["roi"], cuda), _t(c["grid"], cuda), 0.01, 0.0)
    assert p.shape == (0, 2) and r.numel() == 0 and t0.shape == (0, 1)
    # cpu tensors are refused like the reference (NA/ray_marching.py:131-132)
    with pytest.raises(NotImplementedError):
        na.ray_marching(torch.zeros(1, 3), torch.zeros(1, 3), t_min=torch.zeros(1), t_max=torch.ones(1))


def test_pack_info_and_goldens(cuda):
    from supernormal_b200 import nerfacc_api as na
    idx = torch.tensor([0, 2, 2, 2, 2], device=cuda)
    assert na.pack_info(idx, n_rays=3).tolist() == [[0, 1], [1, 0], [1, 4]]   # NT/test_pack.py:27-37
    assert na.pack_info(idx).tolist() == [[0, 1], [1, 0], [1, 4]]
    a = torch.tensor([0.4, 0.3, 0.8, 0.8, 0.5], device=cuda).unsqueeze(-1)
    assert na.render_visibility(a, ray_indices=idx, early_stop_eps=0.03).tolist() == [True, True, True, True, False]
    assert na.render_visibility(a, ray_indices=idx, early_stop_eps=0.05, alpha_thre=0.35).tolist() == [True, False, True, True, False]
    w = na.render_weight_from_alpha(a, ray_indices=idx, n_rays=3)  # NT/test_rendering.py:49-68
    assert torch.allclose(w.cpu(), torch.tensor([0.4, 0.3, 0.56, 0.112, 0.014]).unsqueeze(-1))
    rng = np.random.RandomState(0)
    big = np.sort(rng.randint(0, 5000, 100000))
    assert np.array_equal(na.pack_info(_t(big, cuda)).cpu().numpy(), T.pack_info(torch.from_numpy(big)).numpy())


def test_weight_grads_golden(cuda):  # NT/test_rendering.py:137-214, atol 1e-4 as in the reference
    from supernormal_b200 import nerfacc_api as na
    idx = torch.tensor([0, 2, 2, 2, 2], device=cuda)
    sig = torch.tensor([[0.4], [0.8], [0.1], [0.8], [0.1]], device=cuda, requires_grad=True)
    for kw in (dict(ray_indices=idx, n_rays=3), dict(packed_info=torch.tensor([[0, 1], [1, 0], [1, 4]], dtype=torch.int32, device=cuda))):
        w = na.render_weight_from_alpha(1.0 - torch.exp(-sig), **kw)
        w.sum().backward()
        assert torch.allclose(w.detach().cpu(), torch.tensor([[0.3297], [0.5507], [0.0428], [0.2239], [0.0174]]), atol=1e-4)
        assert torch.allclose(sig.grad.cpu(), torch.tensor([[0.6703], [0.1653], [0.1653], [0.1653], [0.1653]]), atol=1e-4)
        sig.grad = None


@pytest.mark.parametrize("P", [1, 9, 25])
def test_patch_weights_bit_exact(cuda, P):
    from supernormal_b200 import nerfacc_api as na
    rng = np.random.RandomState(P)
    counts = rng.randint(0, 40, size=700)
    counts[::13] = 0
    idx = np.repeat(np.arange(700), counts)
    S = idx.size
    a = rng.uniform(0, 1, (S, P, 1)).astype(np.float32)
    a[rng.randint(0, S, 50)] = 1.0
    a[rng.randint(0, S, 50)] = 0.0
    g = rng.normal(size=(S, P, 1)).astype(np.float32)
    packed = T.pack_info(torch.from_numpy(idx)).numpy()
    w_o = oracle.weight_from_alpha_patch_fwd(packed, a)
    ga_o = oracle.weight_from_alpha_patch_bwd(packed, a, w_o, g)
    at = _t(a, cuda).requires_grad_(True)
    w = na.render_weight_from_alpha_patch_based(at, _t(idx, cuda)) if P > 1 else na.render_weight_from_alpha(at.view(S, 1), ray_indices=_t(idx, cuda)).view(S, 1, 1)
    w.backward(_t(g, cuda))
    assert np.array_equal(w.detach().cpu().numpy().view(np.uint32), w_o.view(np.uint32))
    assert np.array_equal(at.grad.cpu().numpy().view(np.uint32), ga_o.view(np.uint32))


def test_accumulate_fwd_bwd(cuda):
    from supernormal_b200 import nerfacc_api as na
    rng = np.random.RandomState(5)
    counts = rng.randint(0, 30, size=300)
    idx = np.repeat(np.arange(300), counts)
    S, P = idx.size, 9
    w = rng.uniform(0, 1, (S, P, 1)).astype(np.float32)
    v = rng.normal(size=(S, P, 3)).astype(np.float32)
    go = rng.normal(size=(310, P, 3)).astype(np.float32)
    wt, vt = torch.from_numpy(w).requires_grad_(True), torch.from_numpy(v).requires_grad_(True)
    ref = T.accumulate_along_rays_patch_based(wt, torch.from_numpy(idx), values=vt, n_patches=310)
    ref.backward(torch.from_numpy(go))
    wc, vc = _t(w, cuda).requires_grad_(True), _t(v, cuda).requires_grad_(True)
    out = na.accumulate_along_rays_patch_based(wc, _t(idx, cuda), values=vc, n_patches=310)
    out.backward(_t(go, cuda))
    # fp32 sums of <= 30 terms in a different order: 1e-5 absolute on O(1) values
    assert torch.allclose(out.detach().cpu(), ref.detach(), atol=1e-5)
    assert torch.allclose(wc.grad.cpu(), wt.grad, atol=1e-5) and torch.allclose(vc.grad.cpu(), vt.grad, atol=1e-6)
    ws = na.accumulate_along_rays_patch_based(wc.detach(), _t(idx, cuda), n_patches=310)
    assert torch.allclose(ws.cpu(), T.accumulate_along_rays_patch_based(wt.detach(), torch.from_numpy(idx), n_patches=310), atol=1e-5)
    # empty input -> zeros (NA/vol_rendering.py:322-324); per-ray flavour, empty ray -> 0 (NT/test_rendering.py:93-110)
    e = na.accumulate_along_rays_patch_based(torch.zeros(0, 9, 1, device=cuda), torch.zeros(0, dtype=torch.long, device=cuda), n_patches=4)
    assert e.shape == (4, 9, 1) and (e == 0).all()
    i5 = torch.tensor([0, 2, 2, 2, 2], device=cuda)
    w5 = torch.tensor([0.4, 0.3, 0.8, 0.8, 0.5], device=cuda).unsqueeze(-1)
    v5 = torch.rand(5, 2, device=cuda)
    r = na.accumulate_along_rays(w5, i5, values=v5, n_rays=3)
    assert r.shape == (3, 2) and torch.allclose(r[0], w5[0] * v5[0]) and (r[1] == 0).all() and torch.allclose(r[2], (w5[1:] * v5[1:]).sum(0))


def test_occupancy_grid_update(cuda):
    from supernormal_b200 import nerfacc_api as na
    torch.manual_seed(0)
    fn = lambda x: torch.sigmoid(-(x.norm(dim=-1, keepdim=True) - 0.5) * 80)
    g = na.OccupancyGrid([-1, -1, -1, 1, 1, 1], 128).to(cuda)
    go = T.OccupancyGrid([-1, -1, -1, 1, 1, 1], 128)
    assert g.binary.shape == (128, 128, 128) and g.roi_aabb.shape == (6,)   # NT/test_grid.py:16-31
    rand = torch.rand(128 ** 3, 3)
    g._update(0, fn, occ_thre=0.1, rand=rand.to(cuda))
    go.update(0, fn, occ_thre=0.1, rand=rand)
    # sigmoid on GPU vs CPU differs by ulps: compare occs to 1e-6 and binary away from the threshold
    assert torch.allclose(g.occs.cpu(), go.occs, atol=1e-6)
    th = min(float(go.occs.mean()), 0.1)
    safe = (go.occs - th).abs() > 1e-5
    assert torch.equal(g.binary.cpu().flatten()[safe], go.binary.flatten()[safe])
    idx = torch.randint(0, 128 ** 3, (50000,)).unique()  # unique: duplicate-index winner is unspecified in the reference
    rand = torch.rand(idx.numel(), 3)
    g._update(300, fn, occ_thre=0.1, rand=rand.to(cuda), indices=idx.to(cuda))
    go.update(300, fn, occ_thre=0.1, rand=rand, indices=idx)
    assert torch.allclose(g.occs.cpu(), go.occs, atol=1e-6)
    g.eval()
    with pytest.raises(RuntimeError):
        g.every_n_step(0, fn)


def test_ray_marching_wrapper_with_alpha_fn(cuda):
    from supernormal_b200 import nerfacc_api as na
    c = make_march_case(seed=11, n_rays=500, step=0.01)
    grid = na.OccupancyGrid([-1, -1, -1, 1, 1, 1], 128).to(cuda)
    grid._binary = _t(c["grid"], cuda)
    o, d = _t(c["rays_o"], cuda), _t(c["rays_d"], cuda)

    def alpha_fn(t0, t1, ridx):
        mid = o[ridx] + d[ridx] * (t0 + t1) * 0.5
        return (torch.sigmoid(-(mid.norm(dim=-1, keepdim=True) - 0.5) * 40) * 0.9).float()
    ridx, t0, t1 = na.ray_marching(o, d, t_min=_t(c["t_min"], cuda), t_max=_t(c["t_max"], cuda), grid=grid,
                                   render_step_size=np.float64(c["step"]), stratified=False, cone_angle=0.0,
                                   early_stop_eps=1e-3, alpha_fn=alpha_fn)
    oc, dc = torch.from_numpy(c["rays_o"]), torch.from_numpy(c["rays_d"])
    alpha_cpu = lambda a0, a1, r: (torch.sigmoid(-((oc[r] + dc[r] * (a0 + a1) * 0.5).norm(dim=-1, keepdim=True) - 0.5) * 40) * 0.9).float()
    r_o, a0, a1 = T.ray_marching(oc, dc, torch.from_numpy(c["t_min"]), torch.from_numpy(c["t_max"]), torch.from_numpy(c["roi"]),
                                 torch.from_numpy(c["grid"]), c["step"], 0.0, alpha_cpu, early_stop_eps=1e-3)
    # the visibility cut T >= eps can flip for samples whose T is within ulps of eps (sigmoid CPU vs GPU)
    assert abs(ridx.numel() - r_o.numel()) <= 3
    if ridx.numel() == r_o.numel():
        assert torch.equal(ridx.cpu(), r_o) and torch.equal(t0.cpu(), a0)
