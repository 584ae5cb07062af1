"""GPU tests of the harness rows around the hot path:
N3  runner.validate_normal_patch_based / eval_mae (exp_runner.py:397-481, 526-577) against the oracle renderer on a small view;
N4  mesh_post on the device: largest edge-connected cluster (exp_runner.py:508-524) against a union-find oracle on a mesh that the
    GPU marching cubes extracted, and visible surface points (exp_runner.py:580-592) traced through the SDF kernel itself;
a13 runner.render_normal_pixel_based (models/renderer.py:278-351) against the oracle's per-ray operators."""
import numpy as np
import pytest
import torch

import oracle
from oracle import torch_ops as T

pytestmark = pytest.mark.gpu

ENC = dict(otype="HashGrid", n_levels=6, n_features_per_level=2, log2_hashmap_size=14, base_resolution=8, per_level_scale=1.6)


def _pair(cuda, n_patches=512, H=48, W=60, n_active=4, variance=0.3):
    """a FusedTrainer and the CPU oracle networks holding the same (perturbed geometric-init) weights and the same occupancy grid"""
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer
    scene = SyntheticScene(n_views=4, H=H, W=W, exclude_views=(1,))
    ds_cpu, ds = SyntheticDataset(scene, device="cpu"), SyntheticDataset(scene, device=cuda)
    conf = dict(DILIGENT_CONF, batch_size=n_patches, encoding=ENC, end_iter=100)
    torch.manual_seed(0)
    osdf = T.SDFNetwork(ENC, 64, 0.6, fp16=True)
    with torch.no_grad():
        osdf.encoding_params.copy_((torch.rand_like(osdf.encoding_params) * 2 - 1) * 0.02)
        osdf.lin0.weight_v[:, 3:].normal_(0, 0.05)     # weight_norm: a larger perturbation would shrink the xyz columns of g v / |v|
    osdf.bindwidth = n_active
    odev = T.SingleVariance(variance)
    orend = T.NeuSRenderer(osdf, odev)
    r = torch.arange(128).float().add(0.5).div(64).sub(1)
    gx, gy, gz = torch.meshgrid(r, r, r, indexing="ij")
    orend.occupancy_grid.binary = (gx ** 2 + gy ** 2 + gz ** 2).sqrt() < 0.75
    tr = FusedTrainer(ds, conf, device=cuda, samples_per_ray_cap=256)
    sd = {"sdf_network_fine": {"encoding.params": osdf.encoding_params.detach(), "lin0.bias": osdf.lin0.bias.detach(),
                               "lin0.weight_g": osdf.lin0.weight_g.detach(), "lin0.weight_v": osdf.lin0.weight_v.detach(),
                               "lin1.bias": osdf.lin1.bias.detach(), "lin1.weight_g": osdf.lin1.weight_g.detach(),
                               "lin1.weight_v": osdf.lin1.weight_v.detach()},
          "variance_network_fine": {"variance": odev.variance.detach()}}
    tr.model.load_reference_state_dict(sd)
    tr.model.n_active = n_active
    tr.grid._binary = orend.occupancy_grid.binary.to(cuda)
    return ds_cpu, ds, orend, tr


def test_validate_normal_patch_based_and_eval_mae_vs_oracle(cuda):
    """The tiled full-image eval render of one view (non-overlapping 3x3 patches, forward only) equals the oracle's
    NeuSRenderer.render over the same patches (mode='eval', un-jittered), and eval_mae reproduces the angular error computed from
    the oracle's normal map (exp_runner.py:526-577)."""
    from supernormal_b200 import runner
    ds_cpu, ds, orend, tr = _pair(cuda)
    idx = 2
    nm = runner.validate_normal_patch_based(tr, idx, eval_patch_size=128, stratified=False)     # 128: several tiles per view
    ny, nx = ds.H // 3, ds.W // 3
    assert nm.shape == (ny * 3, nx * 3, 3)
    # oracle: the same patch centres (Dataset.gen_patches_at, models/dataset_loader.py:177-221)
    cy, cx = torch.meshgrid(torch.arange(ny) * 3 + 1, torch.arange(nx) * 3 + 1, indexing="ij")
    img = torch.full((ny * nx,), idx, dtype=torch.long)
    o, d, pn, vinv, nrm, msk = ds_cpu.patches_at(img, cx.reshape(-1), cy.reshape(-1), 3, 3)
    near, far = ds_cpu.near_far_from_sphere(o[:, 1, 1], d[:, 1, 1])
    orend.sampling_step_size = tr.step_size(tr.iter_step)
    with torch.no_grad():
        out = orend.render(o, d, pn, near, far, vinv, jitter=None)
    ref = out["comp_normal"].view(ny, nx, 3, 3, 3).permute(0, 2, 1, 3, 4).reshape(ny * 3, nx * 3, 3)
    got = nm.cpu()
    assert out["n_samples"] > 2000
    assert (got - ref).abs().max().item() <= 3e-3 * max(1.0, ref.abs().max().item())
    # mean angular error of the view against the dataset normals, both ways
    def mae(n):
        n = n / (1e-10 + n.norm(dim=-1, keepdim=True))
        gt, mask = ds_cpu.normals[idx, :ny * 3, :nx * 3], ds_cpu.masks[idx, :ny * 3, :nx * 3] > 0.5
        return torch.rad2deg(torch.arccos((gt * n).sum(-1).clamp(-1, 1)))[mask]
    assert abs(mae(got).mean().item() - mae(ref).mean().item()) < 0.05
    res = runner.eval_mae(tr)
    assert set(res) == {"mae_allview", "mae_testview"} and 0 < res["mae_allview"] < 90 and 0 < res["mae_testview"] < 90


def test_mesh_post_on_device(cuda):
    """remove_isolated_clusters on CUDA tensors: a mesh from the GPU marching cubes (sphere of the geometric init) plus a small
    far-away component; the kept cluster must be the union-find oracle's largest, re-indexed.  find_visible_points traced through
    snb_sdf_eval lands on the SDF's zero level set."""
    from supernormal_b200 import mesh, mesh_post, runner
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("_mesh_post_cpu_tests", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_mesh_post.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _uf_clusters = mod._uf_clusters
    ds_cpu, ds, orend, tr = _pair(cuda)
    v, t = mesh.extract_geometry(tr.model, ds.object_bbox_min, ds.object_bbox_max, 48, 0.0)
    v, t = np.asarray(v, np.float32), np.asarray(t, np.int32)
    assert len(t) > 1000
    extra_v = np.array([[3, 3, 3], [3.1, 3, 3], [3, 3.1, 3], [3.1, 3.1, 3]], np.float32)
    extra_t = np.array([[0, 1, 2], [1, 3, 2]], np.int32) + len(v)
    v_all, t_all = np.concatenate([v, extra_v]), np.concatenate([extra_t, t])
    vg, tg = mesh_post.remove_isolated_clusters(torch.from_numpy(v_all).to(cuda), torch.from_numpy(t_all).to(cuda))
    assert vg.is_cuda and tg.is_cuda
    ref = _uf_clusters(t_all)
    big = np.bincount(ref).argmax()
    keep = ref == big
    assert tg.shape[0] == int(keep.sum()) and not keep[:2].any()
    assert np.array_equal(vg.cpu().numpy()[tg.cpu().numpy()], v_all[t_all[keep]])
    assert np.array_equal(np.unique(tg.cpu().numpy()), np.arange(vg.shape[0]))
    # the host-array form (what runner.validate_mesh feeds) gives the same mesh
    v2, t2 = mesh_post.remove_isolated_clusters(v_all, t_all)
    assert np.array_equal(t2, tg.cpu().numpy()) and np.array_equal(v2, vg.cpu().numpy())
    # visible points: first SDF zero crossing along every foreground pixel's ray
    pts = torch.from_numpy(runner.find_visible_points(tr)).to(cuda)
    n_fg = int((ds.masks > 0.5).sum())
    assert pts.shape[1] == 3 and 0.9 * n_fg <= pts.shape[0] <= n_fg
    tr.model.prep()
    assert tr.model.sdf(pts).abs().max().item() < 2e-3
    assert (pts.norm(dim=-1) - pts.norm(dim=-1).mean()).abs().max().item() < 0.3      # a (perturbed) sphere


def test_render_normal_pixel_based_vs_oracle(cuda):
    """runner.render_normal_pixel_based (models/renderer.py:278-351: per-ray marching with alpha_thre = 0, early_stop_eps = 1e-3,
    render_weight_from_alpha / accumulate_along_rays, analytic normals from SDFNetwork.gradient) against the oracle's per-ray
    operators on the same rays, un-jittered."""
    from supernormal_b200 import runner
    ds_cpu, ds, orend, tr = _pair(cuda)
    g = torch.Generator().manual_seed(4)
    view = 2
    py = torch.randint(4, ds_cpu.H - 4, (300,), generator=g)
    px = torch.randint(4, ds_cpu.W - 4, (300,), generator=g)
    o, d, *_ = ds_cpu.patches_at(torch.full((300,), view, dtype=torch.long), px, py, 1, 1)
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    near, far = ds_cpu.near_far_from_sphere(o, d)
    step = 0.01
    out = runner.render_normal_pixel_based(tr, o.to(cuda), d.to(cuda), near.to(cuda), far.to(cuda), step_size=step, stratified=False)
    # ---- oracle, statement by statement like models/renderer.py:278-351
    osdf, odev = orend.sdf_network, orend.deviation_network

    def alpha_fn(t0, t1, ridx):
        with torch.no_grad():
            oo, dd = o[ridx.long()], d[ridx.long()]
            ps, pe = oo + dd * t0, oo + dd * t1
            dm = T._diff_mask(t0, t1)
            sdf = osdf(torch.cat([ps, pe[dm].reshape(-1, 3)], 0))
            s0 = sdf[: ps.shape[0]]
            s1 = T._next_start_or_own_end(s0, sdf[ps.shape[0]:], dm)
            return T.neus_alpha(s0, s1, odev.inv_s()).reshape(-1, 1)
    ridx, t0, t1 = T.ray_marching(o, d, near, far, orend.scene_aabb, orend.occupancy_grid.binary, np.float32(step), 0.0, alpha_fn,
                                  early_stop_eps=1e-3, alpha_thre=0.0, jitter=None)
    alpha = alpha_fn(t0, t1, ridx)
    mid = (t0 + t1) / 2.0
    grad = osdf.gradient(o[ridx.long()] + d[ridx.long()] * mid).reshape(-1, 3).detach()
    w = T.render_weight_from_alpha(alpha, ray_indices=ridx, n_rays=300)
    comp = T.accumulate_along_rays(w, ridx, values=grad, n_rays=300)
    depth = T.accumulate_along_rays(w, ridx, values=mid, n_rays=300)
    wsum = T.accumulate_along_rays(w, ridx, values=None, n_rays=300)
    assert ridx.numel() > 1000 and abs(out["n_samples"] - ridx.numel()) <= max(3, 0.002 * ridx.numel())   # T >= 1e-3 cut: ulp-band samples
    assert (out["weight_sum"].cpu() - wsum).abs().max().item() < 2e-3
    assert (out["comp_depth"].cpu() - depth).abs().max().item() < 5e-3
    assert (out["comp_normal"].cpu() - comp).abs().max().item() < 5e-3 * max(1.0, comp.abs().max().item())
