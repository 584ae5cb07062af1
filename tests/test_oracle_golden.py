"""CPU: pins the oracle against every golden vector the reference's own tests / docstrings hold for
this path.  NT = third_parties/nerfacc-0.3.5/nerfacc-0.3.5/tests."""
import numpy as np
import torch

import oracle
from oracle import torch_ops as T
from conftest import make_march_case


def test_pack_info_golden():  # NT/test_pack.py:27-37
    got = T.pack_info(torch.tensor([0, 2, 2, 2, 2]), n_rays=3)
    assert got.tolist() == [[0, 1], [1, 0], [1, 4]] and got.dtype == torch.int32
    assert T.pack_info(torch.tensor([0, 2, 2, 2, 2])).tolist() == [[0, 1], [1, 0], [1, 4]]


def test_visibility_golden():  # NT/test_rendering.py:19-45
    idx = torch.tensor([0, 2, 2, 2, 2])
    a = torch.tensor([0.4, 0.3, 0.8, 0.8, 0.5]).unsqueeze(-1)
    assert T.render_visibility(a, ray_indices=idx, early_stop_eps=0.03).tolist() == [True, True, True, True, False]
    assert T.render_visibility(a, ray_indices=idx, early_stop_eps=0.05, alpha_thre=0.35).tolist() == [True, False, True, True, False]


def test_visibility_docstring():  # NA/vol_rendering.py:721-728
    a = torch.tensor([[0.4], [0.8], [0.1], [0.8], [0.1], [0.0], [0.9]])
    idx = torch.tensor([0, 0, 0, 1, 1, 2, 2])
    Tr = T.render_transmittance_from_alpha(a, ray_indices=idx)
    assert torch.allclose(Tr.squeeze(-1), torch.tensor([1.0, 0.6, 0.12, 1.0, 0.2, 1.0, 1.0]))
    assert T.render_visibility(a, ray_indices=idx, early_stop_eps=0.3, alpha_thre=0.2).tolist() == [True, True, False, True, False, False, True]


def test_weight_from_alpha_golden():  # NT/test_rendering.py:49-68 and NA/vol_rendering.py:567-571
    a = torch.tensor([0.4, 0.3, 0.8, 0.8, 0.5]).unsqueeze(-1)
    w = T.render_weight_from_alpha(a, ray_indices=torch.tensor([0, 2, 2, 2, 2]), n_rays=3)
    assert torch.allclose(w, torch.tensor([0.4, 0.3, 0.7 * 0.8, 0.14 * 0.8, 0.028 * 0.5]).unsqueeze(-1))
    a = torch.tensor([[0.4], [0.8], [0.1], [0.8], [0.1], [0.0], [0.9]])
    w = T.render_weight_from_alpha(a, ray_indices=torch.tensor([0, 0, 0, 1, 1, 2, 2]))
    assert torch.allclose(w.squeeze(-1), torch.tensor([0.4, 0.48, 0.012, 0.8, 0.02, 0.0, 0.9]))


def test_weight_grads_golden():  # NT/test_rendering.py:137-214 (alpha route), atol 1e-4 as in the reference
    idx = torch.tensor([0, 2, 2, 2, 2])
    sig = torch.tensor([[0.4], [0.8], [0.1], [0.8], [0.1]], requires_grad=True)
    alphas = 1.0 - torch.exp(-sig * 1.0)
    w = T.render_weight_from_alpha(alphas, ray_indices=idx, n_rays=3)
    w.sum().backward()
    assert torch.allclose(w.detach(), torch.tensor([[0.3297], [0.5507], [0.0428], [0.2239], [0.0174]]), atol=1e-4)
    assert torch.allclose(sig.grad, torch.tensor([[0.6703], [0.1653], [0.1653], [0.1653], [0.1653]]), atol=1e-4)


def test_patch_weights_reduce_to_per_ray():  # the only pin the reference offers for *_patch_based
    rng = np.random.RandomState(1)
    counts = rng.randint(0, 9, size=40)
    idx = torch.from_numpy(np.repeat(np.arange(40), counts))
    S, P = idx.numel(), 9
    a = torch.from_numpy(rng.uniform(0, 1, (S, P, 1)).astype(np.float32))
    a[rng.randint(0, S, 5)] = 1.0  # opaque samples: exercises the 1e-10 clamp in the backward
    a.requires_grad_(True)
    g = torch.from_numpy(rng.normal(size=(S, P, 1)).astype(np.float32))
    w = T.render_weight_from_alpha_patch_based(a, idx)
    w.backward(g)
    for k in range(P):
        ak = a.detach()[:, k].clone().requires_grad_(True)
        wk = T.render_weight_from_alpha(ak, ray_indices=idx)
        wk.backward(g[:, k])
        assert torch.equal(wk.detach(), w.detach()[:, k])
        assert torch.equal(ak.grad, a.grad[:, k])


def test_accumulate_golden():  # NT/test_rendering.py:93-110
    idx = torch.tensor([0, 2, 2, 2, 2])
    w = torch.tensor([0.4, 0.3, 0.8, 0.8, 0.5]).unsqueeze(-1)
    v = torch.rand(5, 2)
    r = T.accumulate_along_rays(w, idx, values=v, n_rays=3)
    assert r.shape == (3, 2) and torch.allclose(r[0], w[0] * v[0]) and (r[1] == 0).all()
    assert torch.allclose(r[2], (w[1:] * v[1:]).sum(0))


def test_march_samples_inside_aabb():  # NT/test_ray_marching.py:27-48 (all-true grid -> midpoints in AABB)
    c = make_march_case(seed=3, n_rays=64, grid_kind="full", nan_every=0)
    packed, ridx, t0, t1 = oracle.ray_marching(c["rays_o"], c["rays_d"], c["t_min"], c["t_max"], c["roi"], c["grid"], c["step"])
    assert packed[:, 1].sum() == ridx.shape[0] > 0
    mid = c["rays_o"][ridx] + c["rays_d"][ridx] * (t0 + t1) / 2
    assert (mid >= -1).all() and (mid <= 1).all()
    # contiguous intervals with constant step in fully occupied space
    same = ridx[1:] == ridx[:-1]
    assert np.array_equal(t1[:-1][same], t0[1:][same])


def test_march_nan_and_empty():  # NaN near/far -> 0 samples; empty grid -> 0 samples
    c = make_march_case(seed=4, n_rays=68, nan_every=17)
    packed, ridx, t0, t1 = oracle.ray_marching(c["rays_o"], c["rays_d"], c["t_min"], c["t_max"], c["roi"], c["grid"], c["step"])
    nan_rays = np.isnan(c["t_min"])
    assert nan_rays.sum() >= 4 and (packed[nan_rays, 1] == 0).all()
    assert (packed[:, 0] == np.concatenate([[0], np.cumsum(packed[:, 1])[:-1]])).all()
    c = make_march_case(seed=4, n_rays=16, grid_kind="empty")
    assert oracle.ray_marching(c["rays_o"], c["rays_d"], c["t_min"], c["t_max"], c["roi"], c["grid"], c["step"])[1].size == 0


def test_contraction_aabb():  # NT/test_contraction.py:33-42: roi [-1,1] -> x*0.5+0.5 and back
    g = T.OccupancyGrid([-1, -1, -1, 1, 1, 1], 4)
    seen = {}
    g.update(0, lambda x: seen.setdefault("x", x).sum(-1, keepdim=True) * 0, rand=torch.full((64, 3), 0.5))
    assert torch.allclose((seen["x"] * 0.5 + 0.5) * 4 - 0.5, g.grid_coords.float(), atol=1e-6)


def test_hashgrid_spec_matches_survey():  # SURVEY §8 header: level sizes for the shipped conf
    s = oracle.hashgrid_spec()
    sizes = np.diff(s.offsets.astype(np.int64))
    assert sizes[:4].tolist() == [32768, 79512, 175616, 405224] and (sizes[4:] == 524288).all()
    assert s.n_entries == 5936000 and s.n_params == 11872000
    assert oracle.hashgrid_spec(n_levels=16).n_entries == 6984576


def test_hashgrid_c_vs_torch_and_lattice():
    s = oracle.hashgrid_spec(n_levels=6, log2_hashmap_size=12)
    rng = np.random.RandomState(0)
    params = rng.uniform(-1, 1, s.n_params).astype(np.float32)
    x = rng.uniform(-1, 1, (300, 3)).astype(np.float32)
    out_c = oracle.hashgrid_fwd(s, x, params.astype(np.float16)).astype(np.float32)
    out_t = T.hashgrid_encode(torch.from_numpy(x), torch.from_numpy(params), s).numpy()
    assert np.abs(out_c - out_t).max() <= 4e-3  # fp16 accumulation order vs fp32-then-round: few fp16 ulps of O(1) values
    # at a lattice point of level 0 (scale 31: x = (k-0.5)/31) the encoding is exactly that entry
    k = np.array([[3, 5, 7]], np.float32)
    xl = ((k - 0.5) / np.float32(31)).astype(np.float32)
    idx, w = oracle.hashgrid_corners(s, xl)
    j = int(np.argmax(w[0, 0]))
    assert w[0, 0, j] > 0.999
    o = oracle.hashgrid_fwd(s, xl, params.astype(np.float16)).astype(np.float32)
    tbl = params.astype(np.float16).astype(np.float32).reshape(-1, 2)
    assert np.allclose(o[0, :2], tbl[idx[0, 0, j]], atol=2e-3)


def test_hashgrid_bwd_table_is_adjoint():
    s = oracle.hashgrid_spec(n_levels=5, log2_hashmap_size=11)
    rng = np.random.RandomState(2)
    x = rng.uniform(-1, 1, (200, 3)).astype(np.float32)
    dy = rng.normal(size=(200, s.n_output_dims)).astype(np.float32)
    g = oracle.hashgrid_bwd_table(s, x, dy)
    p = torch.from_numpy(rng.uniform(-1, 1, s.n_params)).double().requires_grad_(True)
    out = T.hashgrid_encode(torch.from_numpy(x).double(), p, s, fp16=False)
    (out * torch.from_numpy(dy).double()).sum().backward()
    assert np.allclose(g, p.grad.numpy(), atol=1e-5)
