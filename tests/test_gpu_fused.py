"""GPU parity of the fused training step (march+visibility, patch SDF forward, render fwd/bwd, MLP +
hash-table backward, weight-norm unfolding, Adam) against the PyTorch-CPU oracle of the reference's
renderer / fields / losses.  Tolerances: marching samples bit-exact; fp32 quantities 2e-4 relative
(different summation orders, fp16-accumulated features are bit-identical on both sides by construction
of the oracle's C path but the torch oracle accumulates them in fp32 -> 1e-3 on features)."""
import math
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import torch_ops as T

pytestmark = pytest.mark.gpu

ENC = dict(otype="HashGrid", n_levels=6, n_features_per_level=2, log2_hashmap_size=14, base_resolution=8, per_level_scale=1.6)


ENC16 = dict(otype="HashGrid", n_levels=16, n_features_per_level=2, log2_hashmap_size=12, base_resolution=4, per_level_scale=1.4)
# the BASELINE / config/diligent.conf:80-87 encoding itself: 14 levels, T = 2^19, base 32, scale 2^0.4 (11 872 000 parameters)
ENC_DILIGENT = dict(otype="HashGrid", n_levels=14, n_features_per_level=2, log2_hashmap_size=19, base_resolution=32, per_level_scale=1.3195079107728942)


def _conf(n_patches, grad="dfd", enc=None):
    from supernormal_b200.synthetic import DILIGENT_CONF
    return dict(DILIGENT_CONF, batch_size=n_patches, encoding=enc or ENC, gradient_method=grad, end_iter=100)


def _setup(cuda, n_patches=96, variance=0.3, n_active=4, seed=0, table_scale=0.02, ENC=ENC):
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene
    from supernormal_b200.trainer import FusedTrainer
    ds = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device="cpu")
    conf = _conf(n_patches, enc=ENC)
    torch.manual_seed(seed)
    osdf = T.SDFNetwork(ENC, 64, 0.6, fp16=True)
    with torch.no_grad():
        osdf.encoding_params.copy_((torch.rand_like(osdf.encoding_params) * 2 - 1) * table_scale)
        osdf.lin0.weight_v[:, 3:].normal_(0, 0.3)
        osdf.lin0.weight_g.mul_(1.3)
    osdf.bindwidth = n_active
    odev = T.SingleVariance(variance)
    orend = T.NeuSRenderer(osdf, odev)
    r = torch.arange(128).float().add(0.5).div(64).sub(1)
    gx, gy, gz = torch.meshgrid(r, r, r, indexing="ij")
    orend.occupancy_grid.binary = (gx ** 2 + gy ** 2 + gz ** 2).sqrt() < 0.7

    ds_gpu = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda)
    tr = FusedTrainer(ds_gpu, conf, device=cuda, samples_per_ray_cap=256)
    sd = {"sdf_network_fine": {"encoding.params": osdf.encoding_params.detach(), "lin0.bias": osdf.lin0.bias.detach(),
                               "lin0.weight_g": osdf.lin0.weight_g.detach(), "lin0.weight_v": osdf.lin0.weight_v.detach(),
                               "lin1.bias": osdf.lin1.bias.detach(), "lin1.weight_g": osdf.lin1.weight_g.detach(),
                               "lin1.weight_v": osdf.lin1.weight_v.detach()},
          "variance_network_fine": {"variance": odev.variance.detach()}}
    tr.model.load_reference_state_dict(sd)
    tr.model.n_active = n_active
    tr.grid._binary = orend.occupancy_grid.binary.to(cuda)
    rng = np.random.RandomState(seed)
    batch_cpu = ds.gen_random_patches(n_patches, 3, 3, np_rng=rng)
    return ds, osdf, odev, orend, tr, batch_cpu


def _to_gpu_batch(ds, batch_cpu, cuda):
    o, d, pn, vinv, nrm, msk = batch_cpu
    near, far = ds.near_far_from_sphere(o[:, 1, 1], d[:, 1, 1])
    g = lambda t: t.contiguous().to(cuda)
    return dict(rays_o=g(o[:, 1, 1]), rays_d=g(d.reshape(-1, 9, 3)), plane_n=g(pn), near=g(near), far=g(far),
                v_inv=g(vinv.reshape(-1, 9, 9)), normal_gt=g(nrm.reshape(-1, 9, 3)), mask=g(msk.reshape(-1, 9))), near, far


def test_sdf_eval_matches_oracle(cuda):
    ds, osdf, odev, orend, tr, _ = _setup(cuda)
    x = torch.rand(5000, 3) * 2 - 1
    tr.model.prep()
    got = tr.model.sdf(x.to(cuda)).cpu()
    # exact-feature oracle: C fp16-faithful features + fp64 MLP
    feats = oracle.hashgrid_fwd(osdf.spec, x.numpy(), osdf.encoding_params.detach().numpy().astype(np.float16), n_active=osdf.bindwidth)
    xin = torch.cat([x, torch.from_numpy(feats.astype(np.float32))], 1).double()
    w0 = (osdf.lin0.weight_g * osdf.lin0.weight_v / osdf.lin0.weight_v.norm(dim=1, keepdim=True)).double()
    w1 = (osdf.lin1.weight_g * osdf.lin1.weight_v / osdf.lin1.weight_v.norm(dim=1, keepdim=True)).double()
    h = torch.nn.functional.softplus(xin @ w0.T + osdf.lin0.bias.double(), beta=100)
    ref = h @ w1.T + osdf.lin1.bias.double()
    assert (got.double() - ref).abs().max() < 2e-5 * max(1.0, ref.abs().max().item())
    occ = tr.model.sdf(x.to(cuda), mode=1).cpu()
    assert torch.allclose(occ.double(), torch.sigmoid(-80 * ref), atol=2e-3)


@pytest.mark.parametrize("n_active,enc", [(0, ENC), (1, ENC), (4, ENC), (6, ENC), (13, ENC16), (16, ENC16)])
def test_sdf_eval_grad_matches_oracle(cuda, n_active, enc):
    """snb_sdf_eval_grad (sdf + analytic gradient, one pass, tensor cores) against an fp64 MLP on the C oracle's
    fp16-faithful features and fp32 dy/dx, and against the autograd route of SDFNetwork.gradient
    (models/fields.py:107-119) on the torch oracle.  Tolerances: sdf 2e-5 abs, gradient 2e-5 * max|grad|."""
    ds, osdf, odev, orend, tr, _ = _setup(cuda, n_active=n_active, ENC=enc)
    n = 4133   # not a multiple of the warp / CTA size: ragged tail
    x = torch.rand(n, 3) * 2 - 1
    tr.model.prep()
    sdf, grad = tr.model.sdf_and_gradient(x.to(cuda))
    sdf, grad = sdf.cpu(), grad.cpu()
    assert torch.equal(sdf, tr.model.sdf(x.to(cuda)).cpu())   # same value path as snb_sdf_eval
    feats, dydx = oracle.hashgrid_fwd(osdf.spec, x.numpy(), osdf.encoding_params.detach().numpy().astype(np.float16),
                                      n_active=osdf.bindwidth, want_dy_dx=True)
    w0 = (osdf.lin0.weight_g * osdf.lin0.weight_v / osdf.lin0.weight_v.norm(dim=1, keepdim=True)).double()
    w1 = (osdf.lin1.weight_g * osdf.lin1.weight_v / osdf.lin1.weight_v.norm(dim=1, keepdim=True)).double()
    xin = torch.cat([x, torch.from_numpy(feats.astype(np.float32))], 1).double()
    z = xin @ w0.T + osdf.lin0.bias.double()
    ref_sdf = torch.nn.functional.softplus(z, beta=100) @ w1.T + osdf.lin1.bias.double()
    jac = torch.cat([torch.eye(3, dtype=torch.float64).expand(n, 3, 3), torch.from_numpy(dydx).double()], 1)   # d x_in / d x  [n, 3+2L, 3]
    dz = torch.einsum("hk,nkd->nhd", w0, jac)
    ref_grad = torch.einsum("nh,nhd->nd", torch.sigmoid(100 * z) * w1[0], dz)
    assert (sdf.double() - ref_sdf).abs().max() < 2e-5 * max(1.0, ref_sdf.abs().max().item())
    scale = max(1.0, ref_grad.abs().max().item())
    assert (grad.double() - ref_grad).abs().max() < 2e-5 * scale, (grad.double() - ref_grad).abs().max().item() / scale
    # autograd route (the reference's own): fp32 expression of the same function
    ag = osdf.gradient(x.clone())[:, 0].detach()
    assert (grad - ag).abs().max() < 1e-3 * scale


def test_sdf_eval_grad_argument_errors(cuda):
    import ctypes as C
    from supernormal_b200 import _lib
    ds, osdf, odev, orend, tr, _ = _setup(cuda)
    net = tr.model.net_struct()
    l = _lib.lib()
    assert l.snb_sdf_eval_grad(-1, None, C.byref(net), None, None, None) < 0
    assert l.snb_sdf_eval_grad(8, None, C.byref(net), None, None, None) < 0 and b"null" in l.snb_last_error()
    assert l.snb_sdf_eval_grad(0, None, C.byref(net), None, None, None) == 0     # empty input is a no-op


@pytest.mark.parametrize("variance,cut_active,n_active,enc", [(0.3, False, 4, ENC), (0.75, True, 4, ENC), (0.3, False, 2, ENC), (0.3, False, 6, ENC),
                                                              (0.3, False, 16, ENC16), (0.3, False, 9, ENC16),
                                                              (0.3, False, 3, ENC_DILIGENT), (0.3, False, 14, ENC_DILIGENT), (0.75, True, 14, ENC_DILIGENT)])
def test_fused_forward_backward_vs_oracle(cuda, variance, cut_active, n_active, enc):
    """The backward is the tcgen05 kernel (8-column tiles up to 4 active levels, 32-column tiles beyond; 16 levels is the maximum the
    kernels are compiled for).  test_fma_backward_cross_check reruns two cases on the FMA kernel (SNB_BWD_UMMA=0)."""
    ds, osdf, odev, orend, tr, batch_cpu = _setup(cuda, variance=variance, n_active=n_active, ENC=enc)
    o, d, pn, vinv, nrm, msk = batch_cpu
    batch, near, far = _to_gpu_batch(ds, batch_cpu, cuda)
    step = 0.02
    jitter = torch.rand(o.shape[0])
    orend.sampling_step_size = step
    # ---- oracle
    out = orend.render(o, d, pn, near, far, vinv, jitter=jitter)
    loss, parts = T.losses(out, nrm, msk)
    params = [osdf.encoding_params, osdf.lin0.weight_g, osdf.lin0.weight_v, osdf.lin0.bias, osdf.lin1.weight_g, osdf.lin1.weight_v,
              osdf.lin1.bias, odev.variance]
    loss.backward()
    grads = [p.grad for p in params]
    # ---- fused
    tr.forward_backward(batch, step, jitter.to(cuda))
    lt = tr.loss_terms()
    assert lt["overflow"] == 0
    S_o = out["n_samples"]
    if not cut_active:
        # un-jittered marching + no visibility cut: the sample list must be identical, bit for bit
        assert lt["n_samples"] == S_o
        pidx, t0c, t1c = out["samples"]
        assert torch.equal(tr.buf.patch_idx[:S_o].cpu().long(), pidx)
        assert torch.equal(tr.buf.t0[:S_o].cpu(), t0c[:, 0]) and torch.equal(tr.buf.t1[:S_o].cpu(), t1c[:, 0])
        sdf0 = tr.buf.sdf[:S_o * 9].view(S_o, 9).cpu()
        assert (sdf0 - out["sdf_start"].reshape(S_o, 9)).abs().max() < 2e-5   # identical fp16 features; fp32 MLP summation order only
    else:
        assert S_o < 0.8 * int(orend_full_count(orend, o, d, near, far, jitter, step))  # the cut really removed samples
        assert abs(lt["n_samples"] - S_o) <= max(3, 0.002 * S_o)
    comp = tr.buf.comp.cpu().view(-1, 3, 3, 3)
    wsum = tr.buf.wsum.cpu().view(-1, 3, 3, 1)
    assert (comp - out["comp_normal"]).abs().max() < 3e-3 * max(1.0, out["comp_normal"].abs().max().item())
    assert (wsum - out["weight_sum"]).abs().max() < 1e-3
    for k, tol in (("normal", 3e-3), ("mask", 1e-3), ("eikonal", 3e-3)):
        assert abs(lt[k] - float(parts[k])) <= tol * max(1.0, abs(float(parts[k]))), (k, lt[k], float(parts[k]))
    # ---- gradients ------------------------------------------------------------------------------------
    m = tr.model
    off = m._small_offsets()

    def flat_grads():
        g = m.grad.cpu()
        return {"table": g[2560:], "g0": g[off["g0"]:off["b0"]], "v0": g[off["v0"]:off["g0"]].view(64, m.d_in), "b0": g[off["b0"]:off["v1"]],
                "g1": g[off["g1"]], "v1": g[off["v1"]:off["g1"]], "var": g[off["var"]]}
    ref = {"table": grads[0], "g0": grads[1].flatten(), "v0": grads[2], "b0": grads[3], "g1": grads[4].flatten()[0],
           "v1": grads[5].flatten(), "var": grads[7]}
    # (C) end to end.  alpha = clip(raw, 0, 1) has a discontinuous derivative at raw == 0; a sample whose raw is within
    # fp32 rounding of 0 may pass its gradient on one side only (measured: 1 of 2e4 samples, +-0.02 on d_sdf), hence 2e-2.
    got = flat_grads()
    for k in ref:
        relk = (got[k] - ref[k]).norm().item() / max(ref[k].norm().item(), 1e-12)
        assert relk <= 2e-2, (k, relk)
    if cut_active:
        return
    # (A) render backward in isolation: d loss / d sdf, excluding samples within 1e-6 of the clip boundary (+ linked neighbours)
    S = S_o
    g_o = out["sdf_all"].grad.reshape(-1, 9)
    inv_s = float(odev.inv_s())
    c, n = torch.sigmoid(out["sdf_start"] * inv_s), torch.sigmoid(out["sdf_end"] * inv_s)
    raw = ((c - n + 1e-5) / (c + 1e-5)).detach().reshape(S, 9)
    # the SDF agrees to ~4e-7, so sigmoid(inv_s * sdf) agrees to ~2e-6: a clip decision can flip when the numerator
    # (raw = num / den) is that close to the boundary
    num, den = (c - n + 1e-5).detach().reshape(S, 9), (c + 1e-5).detach().reshape(S, 9)
    cd, nd = c.detach().reshape(S, 9), n.detach().reshape(S, 9)
    band = (cd * (1 - cd) + nd * (1 - nd)) * inv_s * 1e-6 + 3e-7   # |d sigmoid| for |d sdf| <= 1e-6, plus fp32 rounding of c - n
    risky = (num.abs() < band) | ((num - den).abs() < band)
    risky[1:] |= risky[:-1].clone()
    d0, d1 = tr.buf.d_sdf0[:9 * S].view(S, 9).cpu(), tr.buf.d_sdf1[:9 * S].view(S, 9).cpu()
    es = tr.buf.end_slot[:S].cpu()
    assert torch.equal(out["diff_mask"], es >= 0)
    g_start = d0.clone()
    link = es[:-1] < 0
    g_start[1:][link] += d1[:-1][link]
    g_end = d1[es >= 0]
    scale = g_o.abs().max().item()
    assert ((g_start - g_o[:S]).abs()[~risky]).max().item() <= 2e-4 * scale
    assert ((g_end - g_o[S:]).abs()[~risky[es >= 0]]).max().item() <= 2e-4 * scale
    assert risky.float().mean() < 0.03
    # (B) MLP + hash-table backward in isolation: feed the oracle's seeds, compare parameter gradients tightly
    import ctypes as C
    from supernormal_b200._lib import call, ptr
    from supernormal_b200.trainer import make_batch_struct, SMALL_PAD
    tr.buf.d_sdf0[:9 * S] = g_o[:S].flatten().to(cuda)
    seed1 = torch.zeros(S, 9)
    seed1[es >= 0] = g_o[S:]
    tr.buf.d_sdf1[:9 * S] = seed1.flatten().to(cuda)
    tr.buf.end_slot[:S] = torch.where(es >= 0, es, torch.full_like(es, 1 << 30)).to(cuda)  # no link terms: they are inside g_o already
    m.grad.zero_()
    m.net_grad.zero_()
    bs = make_batch_struct(batch["rays_o"], batch["rays_d"], batch["plane_n"], batch["near"], batch["far"], batch["v_inv"], batch["normal_gt"], batch["mask"])
    net = m.net_struct()
    # what the trainer launches: beyond 4 live levels the split form (tcgen05 MLP backward -> workspace -> scatter kernel), else one kernel
    call("snb_sdf_bwd_patch_ws", C.byref(bs), C.byref(net), C.byref(tr.buf.struct), ptr(tr.buf.feats), ptr(tr.buf.d_sdf0), ptr(tr.buf.d_sdf1),
         ptr(m.grad[SMALL_PAD:]), ptr(m.net_grad), ptr(tr.buf.bwd_ws), tr.buf.bwd_ws_bytes)
    tr.buf.stats[4] = 0.0
    call("snb_unfold_grads", m.n_levels, ptr(m.small), ptr(m.net_grad), ptr(tr.buf.stats), ptr(m.grad))
    got = flat_grads()
    if n_active > 4:
        # cross-check of the two forms on identical inputs: the single fused kernel (no workspace) must give the same gradients up to
        # fp32 summation order (the table sums the same TF32-rounded products in a different order; the MLP part is the same code)
        split_table, split_net = m.grad[SMALL_PAD:].clone(), m.net_grad.clone()
        m.grad.zero_()
        m.net_grad.zero_()
        call("snb_sdf_bwd_patch", C.byref(bs), C.byref(net), C.byref(tr.buf.struct), ptr(tr.buf.feats), ptr(tr.buf.d_sdf0), ptr(tr.buf.d_sdf1),
             ptr(m.grad[SMALL_PAD:]), ptr(m.net_grad))
        fused_table, fused_net = m.grad[SMALL_PAD:], m.net_grad
        assert (split_table - fused_table).abs().max().item() <= 2e-5 * fused_table.abs().max().item()
        assert (split_net - fused_net).abs().max().item() <= 2e-5 * fused_net.abs().max().item()
        assert ((split_table == 0) != (fused_table == 0)).sum().item() <= 8    # the same entries are touched (up to exact cancellations)
    # The tcgen05 backward forms d loss/d features (-> table) and dW0 / db0 from dz rounded to TF32 (round to
    # nearest: unbiased, 2^-12 rms relative per element -- the precision class of the fp16 dL/dy tiny-cuda-nn's own backward
    # consumes); sums over 64 hidden units / many points average that down.  z recompute, dW1, db1 are fp32-accurate.
    for k, tol, tol_max in (("table", 6e-4, 2e-3), ("g0", 5e-4, 1e-3), ("v0", 5e-4, 1e-3), ("b0", 5e-4, 1e-3), ("g1", 2e-4, 5e-4), ("v1", 2e-4, 5e-4)):
        relk = (got[k] - ref[k]).norm().item() / max(ref[k].norm().item(), 1e-12)
        assert relk <= tol, (k, relk)
        assert (got[k] - ref[k]).abs().max().item() <= tol_max * max(ref[k].abs().max().item(), 1e-8), k


def orend_full_count(orend, o, d, near, far, jitter, step):
    ridx, _, _ = T.ray_marching(o[:, 1, 1], d[:, 1, 1], near, far, orend.scene_aabb, orend.occupancy_grid.binary, np.float32(step), 0.0, None, jitter=jitter)
    return ridx.numel()


def test_adam_and_schedule_vs_torch(cuda):
    from supernormal_b200._lib import call, ptr
    n = 100003
    torch.manual_seed(0)
    p = torch.randn(n + 1)[:n].clone()
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-4)
    pc, m, v = p.to(cuda).contiguous(), torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    p16 = torch.zeros(n, dtype=torch.float16, device=cuda)
    for t in range(1, 6):
        gr = torch.randn(n) * (t % 2)   # zero-gradient steps keep moving the parameters through the momentum
        ref.grad = gr.clone()
        opt.step()
        gc = gr.to(cuda)
        call("snb_adam_step", n, ptr(pc), ptr(gc), ptr(m), ptr(v), ptr(p16), 5e-4, 0.9, 0.999, 1e-8, t, 1.0)
        assert (gc == 0).all()
        assert torch.allclose(pc.cpu(), ref.detach(), rtol=1e-5, atol=1e-7)
    assert torch.equal(p16.cpu(), pc.cpu().half())


def test_training_reduces_loss_and_checkpoint_roundtrip(cuda):
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer
    ds = SyntheticDataset(SyntheticScene(n_views=8, H=96, W=128, exclude_views=(0,)), device=cuda)
    conf = dict(DILIGENT_CONF, batch_size=512, end_iter=300, increase_bindwidth_every=30, warm_up_end=10)
    tr = FusedTrainer(ds, conf, device=cuda)
    losses = []
    for it in range(120):
        tr.train_step()
        if it % 20 == 19:
            losses.append(tr.loss_terms())
    assert all(l["overflow"] == 0 for l in losses)
    assert all(math.isfinite(l["loss"]) for l in losses)
    assert losses[-1]["normal"] < 0.5 * losses[0]["normal"] + 1e-3, losses
    assert tr.model.n_active == 4 and tr.iter_step == 120
    # SDF of the sphere r=0.5: zero level set roughly at radius 0.5 after training
    dirs = torch.nn.functional.normalize(torch.randn(2000, 3, device=cuda), dim=-1)
    tr.model.prep()
    inside, outside = tr.model.sdf(dirs * 0.3), tr.model.sdf(dirs * 0.8)
    assert (inside < 0).float().mean() > 0.95 and (outside > 0).float().mean() > 0.95
    sd = tr.model.reference_state_dict()
    assert set(sd["sdf_network_fine"]) == {"encoding.params", "lin0.bias", "lin0.weight_g", "lin0.weight_v", "lin1.bias", "lin1.weight_g", "lin1.weight_v"}
    assert sd["sdf_network_fine"]["encoding.params"].shape == (11872000,) and sd["sdf_network_fine"]["lin0.weight_v"].shape == (64, 31)
    before = tr.model.sdf(dirs * 0.5).clone()
    tr.model.flat.zero_()
    tr.model.load_reference_state_dict(sd)
    tr.model.prep()
    assert torch.equal(tr.model.sdf(dirs * 0.5), before)


def test_device_sampler_matches_dataset(cuda):
    """snb_sample_patches vs the ATen-op restatement of Dataset.gen_random_patches (models/dataset_loader.py:223-297)."""
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer
    ds = SyntheticDataset(SyntheticScene(n_views=8, H=96, W=128, exclude_views=(0, 4)), device=cuda)
    tr = FusedTrainer(ds, dict(DILIGENT_CONF, batch_size=4096), device=cuda)
    b, jit = tr.sample_batch_device(3)
    # recover (view, cx, cy) of every patch from its camera centre and centre ray
    cam = ds.pose_all[:, :3, 3]
    view = (b["rays_o"][:, None, :] - cam[None]).norm(dim=-1).argmin(1)
    assert set(view.tolist()) <= set(ds.train_images) and len(set(view.tolist())) == len(ds.train_images)
    R = ds.pose_all[view, :3, :3]
    dcam = torch.einsum("nji,nj->ni", R, b["rays_d"][:, 4])
    K = ds.intrinsics_all[view, :3, :3]
    pix = torch.einsum("nij,nj->ni", K, dcam / dcam[:, 2:3])
    cx, cy = pix[:, 0].round().long(), pix[:, 1].round().long()
    assert (pix[:, 0] - cx).abs().max() < 1e-2 and (pix[:, 1] - cy).abs().max() < 1e-2
    assert cx.min() >= 1 and cx.max() <= ds.W - 3 and cy.min() >= 1 and cy.max() <= ds.H - 3   # randint(1, W-2)
    assert cx.min() == 1 and cx.max() == ds.W - 3 and cy.min() == 1 and cy.max() == ds.H - 3    # whole range is reached
    o, d, pn, vinv, nrm, msk = ds.patches_at(view, cx, cy)
    near, far = ds.near_far_from_sphere(o[:, 1, 1], d[:, 1, 1])
    assert torch.equal(b["normal_gt"], nrm.view(-1, 9, 3)) and torch.equal(b["mask"], msk.view(-1, 9))
    # V_inverse: the sampler's closed form (R A^-1, no table) against torch.inverse of [ray; right; down] (models/dataset_loader.py:126-137)
    assert not tr.v_inverse_table
    vref = vinv.view(-1, 9, 9)
    assert (b["v_inv"] - vref).abs().max().item() <= 2e-5 * vref.abs().max().item()
    eye = torch.einsum("nkab,nkbc->nkac", b["v_inv"].view(-1, 9, 3, 3),
                       torch.stack([d.view(-1, 9, 3), R[:, None, :, 0].expand(-1, 9, 3), R[:, None, :, 1].expand(-1, 9, 3)], dim=-2))
    assert (eye - torch.eye(3, device=cuda)).abs().max().item() < 1e-5       # V^-1 V = I on the sampler's own ray directions
    # ... and the table path (SNB_VINV_TABLE=1: Dataset.V_inverse_all gathered per pixel) gives the table's bits
    import ctypes as C
    from supernormal_b200._lib import call
    from supernormal_b200.trainer import SnbDataset
    vtab = ds.V_inverse_all.to(cuda, torch.float32).contiguous()
    ds_tab = SnbDataset(ds.n_images, ds.H, ds.W, len(ds.train_images), *[t.data_ptr() for t in tr._ds_tensors[:4]], vtab.data_ptr(), tr.train_ids.data_ptr())
    call("snb_sample_patches", C.byref(ds_tab), tr.n_patches, tr.seed, 3, C.byref(tr._out_structs[tr._slot]))
    assert torch.equal(tr.own_batch["v_inv"], vref) and torch.equal(tr.own_batch["rays_d"], b["rays_d"])
    tr.sample_batch_device(3)
    assert torch.allclose(b["rays_d"], d.view(-1, 9, 3), atol=2e-6) and torch.equal(b["rays_o"], o[:, 1, 1]) and torch.equal(b["plane_n"], pn)
    ok = ~torch.isnan(near)
    assert torch.equal(torch.isnan(b["near"]), ~ok) and torch.allclose(b["near"][ok], near[ok], atol=1e-5) and torch.allclose(b["far"][ok], far[ok], atol=1e-5)
    assert jit.min() >= 0 and jit.max() < 1 and abs(jit.mean().item() - 0.5) < 0.03
    b2 = {k: v.clone() for k, v in b.items()}
    tr.sample_batch_device(3)
    same = lambda x, y: torch.equal(torch.nan_to_num(x, nan=-7.0), torch.nan_to_num(y, nan=-7.0))
    assert all(same(b2[k], tr.own_batch[k]) for k in b2)                   # counter-based: same (seed, step) -> same batch
    tr.sample_batch_device(4)
    assert not torch.equal(b2["rays_d"], tr.own_batch["rays_d"])


def test_fused_host_step_equals_per_kernel_path(cuda):
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer
    ds = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda)
    conf = dict(DILIGENT_CONF, batch_size=300, end_iter=200, increase_bindwidth_every=5)
    a, b = FusedTrainer(ds, conf, device=cuda), FusedTrainer(ds, conf, device=cuda)
    b.fused_host = False
    b.legacy_render = True
    for it in range(12):
        a.train_step()
        # same batch / jitter / grid for the per-kernel path
        b.grid._binary.copy_(a.grid._binary)
        b.grid.occs.copy_(a.grid.occs)
        b.update_occupancy = lambda it: None
        b.train_step(batch={k: v.clone() for k, v in a.own_batch.items()}, jitter=a.own_jitter.clone())
        if it == 0:   # identical inputs and parameters: identical results (the forward has no atomics)
            assert a.buf.totals.tolist() == b.buf.totals.tolist()
            # a: single fused render kernel (warp scans), b: render_fwd/patch_loss/render_bwd (serial chains): fp32 order only
            assert torch.allclose(a.buf.comp, b.buf.comp, atol=2e-5, rtol=1e-4) and torch.allclose(a.buf.wsum, b.buf.wsum, atol=2e-6)
            assert torch.allclose(a.buf.stats[:4], b.buf.stats[:4], rtol=2e-4, atol=1e-6)
            assert torch.allclose(a.buf.stats[4], b.buf.stats[4], rtol=5e-2, atol=1e-6)   # d inv_s: a heavily cancelling sum
            nS = 9 * a.buf.totals[0].item()
            for x, y in ((a.buf.d_sdf0[:nS], b.buf.d_sdf0[:nS]), (a.buf.d_sdf1[:nS], b.buf.d_sdf1[:nS])):
                # dalpha = (gw*T - A) / (1 - alpha):  gw*T - A collapses to ~gw*T_end on rays whose seed is dominated by the
                # constant opacity term (background-masked pixel on the object), so both summation orders -- the reference's
                # serial one and the scan -- carry relative noise ~1e-7 / T_end there.  Elementwise: small absolute OR 2 % relative.
                tol = 2e-4 * y.abs().max().item() + 2e-2 * y.abs()
                assert ((x - y).abs() <= tol).float().mean().item() > 0.995
                assert (x - y).norm().item() <= 0.1 * y.norm().item()
            # one Adam step moves every touched parameter by ~lr*sign(g); only entries whose gradient is rounding noise
            # (fp32 atomic order) may disagree
            assert ((a.model.flat - b.model.flat).abs() > 1e-6).float().mean().item() < 1e-3
        # afterwards Adam's normalisation amplifies that noise on near-zero-gradient entries: compare behaviour, not bits
        la, lb = a.loss_terms(), b.loss_terms()
        assert abs(la["n_samples"] - lb["n_samples"]) <= 0.02 * la["n_samples"] + 5
        assert abs(la["loss"] - lb["loss"]) <= 0.05 * abs(la["loss"]) + 1e-3
    assert a.model.n_active == 3


def test_fused_occupancy_update(cuda):
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer
    ds = SyntheticDataset(SyntheticScene(n_views=4, H=64, W=80, exclude_views=(0,)), device=cuda)
    a, b = FusedTrainer(ds, dict(DILIGENT_CONF, batch_size=64), device=cuda), FusedTrainer(ds, dict(DILIGENT_CONF, batch_size=64), device=cuda)
    b.fused_host = False
    for it in (0, 8, 256, 264):   # two warm-up sweeps (all cells), two sparse updates (uniform + occupied cells)
        a.update_occupancy(it)
        b.update_occupancy(it)
        # geometric-init SDF ~ |x| - 0.6: occupied <=> inside r ~ 0.63; jitter differs, so compare away from the surface
        agree = (a.grid.binary == b.grid.binary).float().mean().item()
        assert agree > 0.985, (it, agree)
        frac = a.grid.binary.float().mean().item()
        assert 0.08 < frac < 0.25, frac
    # ground truth from the SDF itself at the cell centres (jitter moves a sample by at most half a cell diagonal ~0.014)
    r = torch.arange(128, device=cuda).float().add(0.5).div(64).sub(1)
    gx, gy, gz = torch.meshgrid(r, r, r, indexing="ij")
    a.model.prep()
    sdf = a.model.sdf(torch.stack([gx, gy, gz], -1).reshape(-1, 3)).view(128, 128, 128)
    assert a.grid.binary[sdf < -0.03].all() and not a.grid.binary[sdf > 0.08].any()


def test_host_batch_feeder_matches_device_batches(cuda):
    """HostBatchFeeder (pinned staging, one H2D copy per step on a copy stream, async loss read-back) trains exactly like
    handing the same batches over as device tensors; the logged losses are the steps' loss terms."""
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene
    from supernormal_b200.trainer import FusedTrainer
    ds = SyntheticDataset(SyntheticScene(n_views=4, H=64, W=80, exclude_views=(0,)), device=cuda)
    conf = _conf(128)
    a, b = FusedTrainer(ds, conf, device=cuda, seed=0), FusedTrainer(ds, conf, device=cuda, seed=0)
    host = [{k: v.cpu() for k, v in a.sample_batch().items()} for _ in range(5)]
    jit = [torch.rand(128) for _ in range(5)]
    feeder = b.host_feeder(depth=2, log_capacity=8)
    assert feeder.h2d_bytes >= sum(v.numel() * 4 for v in host[0].values()) + 128 * 4 and feeder.d2h_bytes == 48
    ref_losses = []
    feeder.submit(host[0], jit[0])                   # unpacked dict: packed into the staging slot by submit()
    for i in range(5):
        a.train_step(batch={k: v.to(cuda) for k, v in host[i].items()}, jitter=jit[i].to(cuda))
        ref_losses.append(a.loss_terms())
        if i + 1 < 5:
            if i % 2:
                feeder.submit(host[i + 1], jit[i + 1])
            else:
                feeder.submit(feeder.pack(host[i + 1], jit[i + 1]))   # pre-packed pinned buffer: one memcpy
        feeder.step()
    got = feeder.losses()
    assert len(got) == 5
    # the two runs see identical inputs; fp32 atomics reorder run to run and Adam's m/sqrt(v) amplifies that on parameters
    # with near-zero gradients, so trajectories agree closely but not bit for bit
    assert got[0]["n_samples"] == ref_losses[0]["n_samples"] and abs(got[0]["loss"] - ref_losses[0]["loss"]) <= 1e-6 * max(1.0, abs(ref_losses[0]["loss"]))
    for r, g in zip(ref_losses, got):
        assert g["overflow"] == 0 and abs(g["n_samples"] - r["n_samples"]) <= max(3, 0.01 * r["n_samples"])
        assert abs(g["loss"] - r["loss"]) <= 5e-3 * max(1.0, abs(r["loss"]))
    assert (a.model.flat - b.model.flat).abs().max() <= 5 * 5e-4 * 2   # bounded by steps * lr per parameter


def test_fma_backward_cross_check(cuda):
    """The thread-per-point FMA backward (the pre-tcgen05 kernel, kept as the independent implementation) still passes the oracle
    parity cases: the kernel choice is read once per process from SNB_BWD_UMMA, hence the subprocess."""
    import os, subprocess, sys
    env = dict(os.environ, SNB_BWD_UMMA="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-k",
                        "test_fused_forward_backward_vs_oracle and (0.3-False-2 or 0.3-False-6)"], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "2 passed" in r.stdout, r.stdout[-2000:]


def test_train_tail_bit_identical_to_separate_kernels(cuda):
    """snb_train_tail -- weight-norm backward, Adam (MLP block + live table levels + fp16 refresh), fold of the updated weights,
    net_grad reset and the next batch, ONE launch -- against snb_unfold_grads -> snb_train_optim -> snb_prep_net ->
    snb_sample_patches on the same state: every buffer bit-identical, with and without pre-unfolded gradients."""
    import ctypes as C
    from supernormal_b200._lib import call, ptr
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer
    ds = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda)
    tr = FusedTrainer(ds, dict(DILIGENT_CONF, batch_size=300, end_iter=200, increase_bindwidth_every=1), device=cuda)
    for _ in range(5):   # non-trivial parameters and Adam state, 5 live levels
        tr.train_step()
    m = tr.model
    batch, jitter = tr.sample_batch_device(5)
    tr.forward_backward(batch, tr.step_size(5), jitter, lean=True)   # gradient w.r.t. the folded weights in net_grad, table gradient in grad
    names = ("flat", "grad", "exp_avg", "exp_avg_sq", "table_f16", "net", "net_grad")
    snap = {k: getattr(m, k).clone() for k in names}
    stats = tr.buf.stats.clone()
    assert snap["net_grad"].abs().sum() > 0 and snap["grad"].abs().sum() > 0 and stats[4] != 0
    t, lr = 6, 3e-4
    nxt = 1 - tr._slot

    def restore():
        for k in names:
            getattr(m, k).copy_(snap[k])
        tr.buf.stats.copy_(stats)
        for v in tr._own[nxt].values():
            v.fill_(-7.0)

    def result():
        torch.cuda.synchronize()
        return {**{k: getattr(m, k).clone() for k in names}, **{"b_" + k: v.clone() for k, v in tr._own[nxt].items()},
                "jit": tr._own_jitter[nxt].clone()}

    restore()
    ctx = tr._ctx(batch, None)
    call("snb_unfold_grads", m.n_levels, ptr(m.small), ptr(m.net_grad), ptr(tr.buf.stats), ptr(m.grad))
    call("snb_train_optim", C.byref(ctx), lr, t, 1.0)
    m.prep()
    m.net_grad.zero_()
    call("snb_sample_patches", C.byref(tr.ds_struct), tr.n_patches, tr.seed, 6, C.byref(tr._out_structs[nxt]))
    ref = result()
    assert not torch.equal(ref["flat"], snap["flat"]) and not torch.equal(ref["net"], snap["net"])

    restore()
    call("snb_train_tail", C.byref(ctx), lr, t, 1.0, 0, C.byref(tr.ds_struct), tr.n_patches, tr.seed, 6, C.byref(tr._out_structs[nxt]))
    got = result()
    for k in ref:   # near/far are NaN for rays that miss the unit sphere
        same = torch.equal(torch.nan_to_num(got[k].float(), nan=-123.0), torch.nan_to_num(ref[k].float(), nan=-123.0))
        assert same, (k, (got[k].float() - ref[k].float()).abs().max().item())
    assert (got["grad"] == 0).all() and (got["net_grad"] == 0).all()

    restore()   # data-parallel order: unfold -> (allreduce) -> tail with grads_unfolded = 1, no sampler blocks
    call("snb_unfold_grads", m.n_levels, ptr(m.small), ptr(m.net_grad), ptr(tr.buf.stats), ptr(m.grad))
    call("snb_train_tail", C.byref(ctx), lr, t, 1.0, 1, None, 0, 0, 0, None)
    got = result()
    for k in names:
        assert torch.equal(got[k], ref[k]), k
    assert all((v == -7.0).all() for v in tr._own[nxt].values())


def test_lean_step_equals_separate_launch_step(cuda):
    """train_step with the lean launch sequence (marcher .. backward + ONE tail kernel that also pre-samples the next batch: 6
    launches) against the same step with prep_net / unfold_grads / Adam / sample_patches as separate launches (lean=False)."""
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer
    from supernormal_b200 import _lib
    ds = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda)
    conf = dict(DILIGENT_CONF, batch_size=300, end_iter=200, increase_bindwidth_every=5)
    a, b = FusedTrainer(ds, conf, device=cuda), FusedTrainer(ds, conf, device=cuda)
    assert a.lean
    b.lean = False
    for it in range(12):
        n0 = _lib.LAUNCH_COUNT
        a.train_step()
        if it % 8:   # no occupancy update in this iteration
            assert _lib.LAUNCH_COUNT - n0 == 6   # marcher, compaction, SDF forward, render, SDF backward, tail
        b.grid._binary.copy_(a.grid._binary)
        b.grid.occs.copy_(a.grid.occs)
        b.update_occupancy = lambda it: None
        b.train_step()
        for k in a.own_batch:   # pre-sampled by the previous tail kernel (a) vs sampled at the start of the step (b)
            assert torch.equal(torch.nan_to_num(a.own_batch[k]), torch.nan_to_num(b.own_batch[k])), (it, k)
        assert torch.equal(a.own_jitter, b.own_jitter)
        if it == 0:   # identical inputs and parameters: identical forward; gradients differ by fp32 atomic order only
            assert a.buf.totals.tolist() == b.buf.totals.tolist()
            assert torch.equal(a.buf.stats[0], b.buf.stats[0])
            assert torch.allclose(a.buf.stats[:4], b.buf.stats[:4], rtol=1e-5, atol=1e-7)
            assert ((a.model.flat - b.model.flat).abs() > 1e-6).float().mean().item() < 1e-3
            b.model.prep()   # a's tail kernel already folded the updated weights
            assert torch.allclose(a.model.net, b.model.net, rtol=1e-4, atol=1e-6)
        la, lb = a.loss_terms(), b.loss_terms()
        assert abs(la["n_samples"] - lb["n_samples"]) <= 0.02 * la["n_samples"] + 5
        assert abs(la["loss"] - lb["loss"]) <= 0.05 * abs(la["loss"]) + 1e-3
    assert a.model.n_active == 3 and (a.model.net_grad == 0).all()


@pytest.mark.parametrize("cap", [320, 4])
def test_single_launch_compaction_equals_scan_then_compact(cuda, cap):
    """snb_compact_samples_stats (per-CTA prefix sums + compaction + loss-accumulator reset, one launch) against
    snb_compact_samples (single-CTA scan kernel, then compaction) on the same marcher output -- also when the sample lists
    overflow their capacity (cap = 4 samples per ray) and have to be clipped."""
    import ctypes as C
    from supernormal_b200._lib import call, ptr
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer, make_batch_struct
    ds = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda)
    tr = FusedTrainer(ds, dict(DILIGENT_CONF, batch_size=301), device=cuda, samples_per_ray_cap=cap)   # 301: ragged last CTA
    tr.update_occupancy(0)
    tr.model.n_active = 1
    batch, jitter = tr.sample_batch_device(0)
    b = tr.buf
    bs = make_batch_struct(*[batch[k] for k in ("rays_o", "rays_d", "plane_n", "near", "far", "v_inv", "normal_gt", "mask")])
    net = tr.model.net_struct()
    rs = C.byref(b.struct)
    call("snb_march_visible", C.byref(bs), C.byref(net), ptr(tr.grid.roi_aabb), *tr.grid._res, ptr(tr.grid.binary.view(torch.uint8)), 0.01,
         ptr(jitter), 1e-8, rs)
    outs = ("packed_info", "end_packed", "totals", "t0", "t1", "patch_idx", "end_slot", "slot_sample")
    march_totals = b.totals.clone()

    def run(name, *extra):
        for k in outs:
            getattr(b, k).fill_(-1)
        b.totals.copy_(march_totals)
        b.stats.fill_(3.0)
        call(name, tr.n_patches, rs, *extra)
        torch.cuda.synchronize()
        return {k: getattr(b, k).clone() for k in outs}

    ref = run("snb_compact_samples")
    got = run("snb_compact_samples_stats", batch["mask"].numel(), ptr(batch["mask"]), ptr(b.stats))
    S, E = ref["totals"][0].item(), ref["totals"][1].item()
    assert S > 0 and E > 0 and (ref["totals"][2].item() == 1) == (cap == 4)
    for k in outs:
        assert torch.equal(got[k], ref[k]), k
    assert b.stats[0].item() == pytest.approx((batch["mask"] > 0.5).sum().item() + 1e-5) and (b.stats[1:] == 0).all()
    # launch order of the render stage (snb_samples.launch_order): the one-launch form leaves the patches sorted by size class
    # (chunks of 32 samples, capped at 7; empty patches last), index order within a class; the two-launch form the identity
    cnt = b.counts.long()                                       # un-clipped marcher counts: what the order is built from
    cls = torch.where(cnt <= 0, torch.zeros_like(cnt), torch.clamp((cnt + 31) // 32, max=7))
    want = torch.sort(-cls * (tr.n_patches + 1) + torch.arange(tr.n_patches, device=cuda), stable=True).indices   # class descending, then index
    assert torch.equal(b.launch_order.long(), want)
    call("snb_compact_samples", tr.n_patches, rs)
    assert torch.equal(b.launch_order.long(), torch.arange(tr.n_patches, device=cuda))


def test_render_launch_order_changes_no_output(cuda):
    """snb_render_fused with the launch order left by the compaction (longest patches first) against the identity order: every
    per-patch / per-sample output is bit-identical (a CTA's work does not depend on which CTA it is); the loss sums meet in atomics."""
    import ctypes as C
    from supernormal_b200._lib import call, ptr
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer, make_batch_struct, SnbSamples
    ds = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda)
    tr = FusedTrainer(ds, dict(DILIGENT_CONF, batch_size=301), device=cuda)
    for _ in range(3):
        tr.train_step()
    batch, jitter = tr.sample_batch_device(7)
    tr.forward_backward(batch, 0.01, jitter, lean=True)   # what train_step runs: march + one-launch compaction (+ order) + forward + render + backward
    torch.cuda.synchronize()
    b, c = tr.buf, tr.conf
    order = b.launch_order.clone()
    assert not torch.equal(order.long(), torch.arange(tr.n_patches, device=cuda)) and torch.equal(order.long().sort().values, torch.arange(tr.n_patches, device=cuda))
    bs = make_batch_struct(*[batch[k] for k in ("rays_o", "rays_d", "plane_n", "near", "far", "v_inv", "normal_gt", "mask")])
    net = tr.model.net_struct()
    S = int(b.totals[0])
    outs = {}
    for name, lo in (("sorted", order), ("identity", None)):
        st = SnbSamples.from_buffer_copy(b.struct)
        st.launch_order = lo.data_ptr() if lo is not None else None
        for t in (b.comp, b.wsum, b.d_sdf0, b.d_sdf1):
            t.fill_(7.0)
        b.stats[1:].zero_()
        call("snb_render_fused", C.byref(bs), C.byref(net), C.byref(st), ptr(b.sdf), float(c["normal_weight"]), float(c["mask_weight"]),
             float(c["eikonal_weight"]), ptr(b.comp), ptr(b.wsum), ptr(b.d_sdf0), ptr(b.d_sdf1), ptr(b.stats))
        torch.cuda.synchronize()
        outs[name] = (b.comp.clone(), b.wsum.clone(), b.d_sdf0[: 9 * S].clone(), b.d_sdf1[: 9 * S].clone(), b.stats.clone())
    assert S > 500
    for x, y in zip(outs["sorted"][:4], outs["identity"][:4]):
        assert torch.equal(x, y)
    assert torch.allclose(outs["sorted"][4], outs["identity"][4], rtol=1e-5, atol=1e-6)


def test_checkpoint_resume_continues_training(cuda):
    """FusedTrainer.state_dict / load_state_dict (exp_runner.py:298-315 + the occupancy grid): a fresh trainer resumed from the
    checkpoint holds the same bits and its next iteration sees the same batch, the same samples and the same loss."""
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer
    ds = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda)
    conf = dict(DILIGENT_CONF, batch_size=256, end_iter=200, increase_bindwidth_every=5, warm_up_end=10)
    a = FusedTrainer(ds, conf, device=cuda)
    for _ in range(20):
        a.train_step()
    sd = a.state_dict()
    assert {"sdf_network_fine", "variance_network_fine", "optimizer", "iter_step", "occupancy_grid"} <= set(sd)
    b = FusedTrainer(ds, conf, device=cuda)
    import io
    f = io.BytesIO()
    torch.save(sd, f)                              # exp_runner.py:315
    f.seek(0)
    b.load_state_dict(torch.load(f, map_location="cpu"))   # exp_runner.py:299
    assert b.iter_step == 20 and b.model.n_active == a.model.n_active == 4 and b.lr == pytest.approx(a.lr, rel=1e-12)
    for name in ("flat", "exp_avg", "exp_avg_sq", "table_f16"):
        assert torch.equal(getattr(a.model, name), getattr(b.model, name)), name
    assert torch.equal(a.grid.binary, b.grid.binary) and torch.equal(a.grid.occs, b.grid.occs)
    a.train_step()
    b.train_step()
    for k in a.own_batch:
        assert torch.equal(torch.nan_to_num(a.own_batch[k]), torch.nan_to_num(b.own_batch[k])), k
    assert a.buf.totals.tolist() == b.buf.totals.tolist()          # same parameters, grid and batch: same samples
    la, lb = a.loss_terms(), b.loss_terms()
    assert lb["loss"] == pytest.approx(la["loss"], rel=1e-4) and a.iter_step == b.iter_step == 21


@pytest.mark.parametrize("f16_only,epoch", [(0, 0), (1, 0), (1, 77)])
def test_peer_tail_world1_equals_train_tail(cuda, f16_only, epoch):
    """snb_train_tail_peer with a peer group of ONE rank (plain device memory stands in for the symmetric allocation): the
    reduction over ranks, the broadcast and both in-kernel barriers degenerate to this GPU, so every buffer must come out
    bit-identical to snb_train_tail -- for the validated variant and for the fp16-only / local-zeroing one."""
    import ctypes as C
    from supernormal_b200 import dp
    from supernormal_b200._lib import call
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer, SnbPeerGroup, NET_FLOATS
    ds = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda)
    tr = FusedTrainer(ds, dict(DILIGENT_CONF, batch_size=300, end_iter=200, increase_bindwidth_every=1), device=cuda)
    for _ in range(5):
        tr.train_step()
    m = tr.model
    batch, jitter = tr.sample_batch_device(5)
    tr.forward_backward(batch, tr.step_size(5), jitter, lean=True)
    names = ("flat", "grad", "exp_avg", "exp_avg_sq", "table_f16", "net", "net_grad")
    snap = {k: getattr(m, k).clone() for k in names}
    stats = tr.buf.stats.clone()
    t, lr = 6, 3e-4

    def restore():
        for k in names:
            getattr(m, k).copy_(snap[k])
        tr.buf.stats.copy_(stats)

    restore()
    ctx = tr._ctx(batch, None)
    call("snb_train_tail", C.byref(ctx), lr, t, 1.0, 0, None, 0, 0, 0, None)
    torch.cuda.synchronize()
    ref = {k: getattr(m, k).clone() for k in names}

    restore()
    m.grad[:NET_FLOATS] = snap["net_grad"]        # peer mode: the folded-weight gradient lives in flat_grad[0:2432)
    flags = torch.zeros(dp.PEER_FLAG_WORDS, dtype=torch.int32, device=cuda)
    counter = torch.zeros(1, dtype=torch.int32, device=cuda)
    one = lambda p: (C.c_void_p * dp.MAX_PEERS)(p, *([None] * (dp.MAX_PEERS - 1)))
    pg = SnbPeerGroup(1, 0, one(m.flat.data_ptr()), one(m.grad.data_ptr()), one(m.table_f16.data_ptr()), one(flags.data_ptr()),
                      counter.data_ptr(), f16_only, epoch)
    ctx = tr._ctx(batch, None)
    ctx.net_grad = m.grad.data_ptr()
    call("snb_train_tail_peer", C.byref(ctx), C.byref(pg), lr, t, None, 0, 0, 0, None)
    torch.cuda.synchronize()
    ep = epoch or t      # the barrier epoch: the launch counter when given, else the optimizer step count
    assert flags[16].item() == 0 and counter.item() == 0 and flags[0].item() == ep and flags[dp.MAX_PEERS].item() == ep
    for k in ("flat", "exp_avg", "exp_avg_sq", "table_f16", "net"):
        assert torch.equal(getattr(m, k), ref[k]), k
    assert (m.grad == 0).all()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_peer_tail_two_ranks_replicas_bit_identical(cuda):
    """torchrun x2 of scripts/dp_peer_check.py: with the peer-memory tail (snb_train_tail_peer) every rank holds the same parameter /
    fp16-table / folded-net bits after 1, 2 and 24 steps, the gradient range is zero after the step, and the trajectory equals the
    NCCL all-reduce path's (same batches; fp32 summation order only)."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29677",
           os.path.join(root, "scripts", "dp_peer_check.py"), "--every", "3", "--steps", "40"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["peer_mode"] == [True, False]
    assert all(a and b for a, b in out["replicas_identical_peer_nccl"]) and all(out["replicas_identical_end"])
    assert out["grad_zero_after_step"] and out["frac_params_apart_gt_1e-6"][0] == 0.0
    assert out["loss_max_rel_diff"] < 1e-3 and out["n_active"] >= 8


def test_checkpoint_interoperates_with_the_reference_format(cuda):
    """exp_runner.py:298-315: a checkpoint is {sdf_network_fine, variance_network_fine, optimizer = torch.optim.Adam.state_dict(),
    iter_step}.  (1) FusedTrainer.state_dict() loads into the reference's own objects -- the literal models/fields.py networks and a
    torch.optim.Adam over `list(sdf_network.parameters()) + list(deviation_network.parameters())` (exp_runner.py:86-97) -- with equal
    tensors; (2) a checkpoint written by those objects (no occupancy grid, no n_active) resumes a FusedTrainer: same parameters and Adam
    moments, n_active derived from iter_step, occupancy grid rebuilt from the SDF."""
    from oracle import ref_models as R
    if not R.available():
        pytest.skip("reference models unavailable")
    from supernormal_b200 import nerfacc_api, tcnn_api
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer, SMALL_PAD
    ds = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda)
    conf = dict(DILIGENT_CONF, batch_size=256, end_iter=200, increase_bindwidth_every=5, warm_up_end=10)
    a = FusedTrainer(ds, conf, device=cuda)
    for _ in range(12):
        a.train_step()
    sd = a.state_dict()
    _, fields_mod = R.load(nerfacc_api, tcnn_api)
    kw = dict(d_out=1, d_in=3, d_hidden=64, n_layers=1, skip_in=[-1], bias=0.6, geometric_init=True, weight_norm=True, input_concat=True)
    sdf_net = fields_mod.SDFNetwork(**kw, encoding_config=conf["encoding"]).to(cuda)
    var_net = fields_mod.SingleVarianceNetwork(init_val=0.5).to(cuda)
    params = list(sdf_net.parameters()) + list(var_net.parameters())
    opt = torch.optim.Adam(params, lr=conf["learning_rate"])
    # (1) ours -> reference objects
    sdf_net.load_state_dict(sd["sdf_network_fine"])
    var_net.load_state_dict(sd["variance_network_fine"])
    opt.load_state_dict(sd["optimizer"])
    assert opt.param_groups[0]["lr"] == pytest.approx(a.lr)
    m = a.model
    for p_, (name, off, shape) in zip(params, a._param_slices()):
        n = p_.numel()
        assert tuple(p_.shape) == tuple(shape), name
        assert torch.equal(p_.detach().flatten(), m.flat[off:off + n]), name
        st = opt.state[p_]
        assert float(st["step"]) == 12 and torch.equal(st["exp_avg"].flatten(), m.exp_avg[off:off + n]), name
        assert torch.equal(st["exp_avg_sq"].flatten(), m.exp_avg_sq[off:off + n]), name
    # (2) reference objects -> ours: what Runner.save_checkpoint writes
    ckpt = {"sdf_network_fine": sdf_net.state_dict(), "variance_network_fine": var_net.state_dict(), "optimizer": opt.state_dict(), "iter_step": 12}
    b = FusedTrainer(ds, conf, device=cuda)
    b.load_state_dict(ckpt)
    assert b.iter_step == 12 and b.model.n_active == a.model.n_active == 3 and b.lr == pytest.approx(a.lr)
    for name in ("flat", "exp_avg", "exp_avg_sq", "table_f16"):
        assert torch.equal(getattr(a.model, name), getattr(b.model, name)), name
    # the grid was rebuilt from the SDF (all-cells sweep): it must cover what the trained grid marks occupied, up to the EMA's memory
    assert b.grid.binary.any()
    inter = (a.grid.binary & b.grid.binary).sum().item()
    assert inter >= 0.9 * min(a.grid.binary.sum().item(), b.grid.binary.sum().item())
    b.train_step()
    assert math.isfinite(b.loss_terms()["loss"])


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("n_active,enc", [(4, ENC), (9, ENC16)])
def test_ad_gradient_method_vs_oracle(cuda, monkeypatch, n_active, enc, fused):
    """gradient_method = 'ad' (models/renderer.py:225-226 + SDFNetwork.gradient, models/fields.py:107-119: analytic normals with
    create_graph=True, i.e. a double backward through the MLP and the hash grid) in FusedTrainer against oracle.torch_ops with
    grad = 'ad': same samples, rendered normals, loss terms and parameter gradients; then it trains.
    fused = 1: the fully fused step (snb_sdf_grad_patch -> snb_render_fused_ad -> snb_sdf_bwd_patch_ws + snb_sdf_grad_bwd_patch);
    fused = 0 (SNB_AD_FUSED=0): the autograd route over the drop-in operators, kept as the cross-check."""
    monkeypatch.setenv("SNB_AD_FUSED", str(fused))
    ds, osdf, odev, orend, tr_dfd, batch_cpu = _setup(cuda, n_active=n_active, ENC=enc)
    from supernormal_b200.trainer import FusedTrainer
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene
    tr = FusedTrainer(SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda), _conf(96, grad="ad", enc=enc), device=cuda,
                      samples_per_ray_cap=256)
    assert tr.gradient_method == "ad" and tr.ad_fused == bool(fused) and tr.lean == bool(fused)
    tr.model.flat.copy_(tr_dfd.model.flat)
    tr.model.refresh_table_f16()
    tr.model.n_active = n_active
    tr.grid._binary = tr_dfd.grid._binary.clone()
    o, d, pn, vinv, nrm, msk = batch_cpu
    batch, near, far = _to_gpu_batch(ds, batch_cpu, cuda)
    step = 0.02
    jitter = torch.rand(o.shape[0])
    orend.sampling_step_size = step
    out = orend.render(o, d, pn, near, far, vinv, jitter=jitter, gradient_method="ad")
    loss, parts = T.losses(out, nrm, msk)
    params = [osdf.encoding_params, osdf.lin0.weight_g, osdf.lin0.weight_v, osdf.lin0.bias, osdf.lin1.weight_g, osdf.lin1.weight_v,
              osdf.lin1.bias, odev.variance]
    loss.backward()
    if fused:
        tr.forward_backward(batch, step, jitter.to(cuda), lean=False)      # per-kernel form: gradients unfolded into model.grad
    else:
        tr.forward_backward_ad(batch, step, jitter.to(cuda))
    lt = tr.loss_terms()
    S = out["n_samples"]
    assert lt["n_samples"] == S and lt["overflow"] == 0
    if fused:   # the analytic gradients themselves, point by point
        g_f = tr.buf.grad[:S * 27].view(S, 3, 3, 3).cpu()
        assert (g_f - out["gradients"].detach()).abs().max().item() <= 2e-3 * max(1.0, out["gradients"].abs().max().item())
    comp = tr.buf.comp.cpu().view(-1, 3, 3, 3)
    assert (comp - out["comp_normal"]).abs().max() < 3e-3 * max(1.0, out["comp_normal"].abs().max().item())
    for k, tol in (("normal", 3e-3), ("mask", 1e-3), ("eikonal", 3e-3)):
        assert abs(lt[k] - float(parts[k])) <= tol * max(1.0, abs(float(parts[k]))), (k, lt[k], float(parts[k]))
    m = tr.model
    off = m._small_offsets()
    g = m.grad.cpu()
    got = {"table": g[2560:], "g0": g[off["g0"]:off["b0"]], "v0": g[off["v0"]:off["g0"]].view(64, m.d_in), "b0": g[off["b0"]:off["v1"]],
           "g1": g[off["g1"]], "v1": g[off["v1"]:off["g1"]], "var": g[off["var"]]}
    ref = {"table": params[0].grad, "g0": params[1].grad.flatten(), "v0": params[2].grad, "b0": params[3].grad, "g1": params[4].grad.flatten()[0],
           "v1": params[5].grad.flatten(), "var": params[7].grad}
    for k in ref:   # the oracle's features are fp16-rounded like ours; its double backward runs in fp32 torch ops on the same function
        relk = (got[k] - ref[k]).norm().item() / max(ref[k].norm().item(), 1e-12)
        assert relk <= 2e-2, (k, relk)
    # ... and the whole step runs: sampler -> ad forward/backward -> fused Adam, loss finite and parameters moving
    before = m.flat.clone()
    tr.iter_step = 20      # past the lr = 0 first step
    tr.lr = 5e-4
    for _ in range(3):
        tr.train_step()
    assert math.isfinite(tr.loss_terms()["loss"]) and not torch.equal(before, m.flat)


def test_ad_fused_step_equals_autograd_route(cuda, monkeypatch):
    """The fused 'ad' step against the autograd route on the SAME trainer state and batch at the diligent encoding (14 levels: split
    backward + gradient-path backward): all parameter gradients agree to fp32 / TF32 accuracy."""
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer
    ds = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda)
    conf = dict(DILIGENT_CONF, batch_size=192, end_iter=200, increase_bindwidth_every=1, gradient_method="ad")
    grads = []
    for fused in (1, 0):
        monkeypatch.setenv("SNB_AD_FUSED", str(fused))
        tr = FusedTrainer(ds, conf, device=cuda)
        g = torch.Generator(device=cuda).manual_seed(5)
        with torch.no_grad():   # make the features matter
            tr.model.table.copy_((torch.rand(tr.model.n_table, device=cuda, generator=g) * 2 - 1) * 0.02)
            o = tr.model._small_offsets()
            tr.model.small[o["v0"]:o["g0"]].view(64, tr.model.d_in)[:, 3:] = torch.randn(64, tr.model.d_in - 3, device=cuda, generator=g) * 0.05
        tr.model.refresh_table_f16()
        tr.model.n_active = 14
        tr.update_occupancy(0)
        batch, jitter = tr.sample_batch_device(3)
        if fused:
            tr.forward_backward(batch, 0.01, jitter, lean=False)
        else:
            tr.forward_backward_ad(batch, 0.01, jitter)
        lt = tr.loss_terms()
        grads.append((tr.model.grad.clone(), lt))
    (ga, la), (gb, lb) = grads
    assert la["n_samples"] == lb["n_samples"] > 1000
    for k in ("normal", "mask", "eikonal"):
        assert abs(la[k] - lb[k]) <= 2e-3 * max(1.0, abs(lb[k])), k
    nt = ga.numel() - 2560
    for name, a, b_ in (("mlp", ga[:2560], gb[:2560]), ("table", ga[2560:], gb[2560:])):
        rel = (a - b_).norm().item() / max(b_.norm().item(), 1e-12)
        assert rel <= 5e-3, (name, rel)


@pytest.mark.parametrize("n_active", [2, 7])
def test_step_with_no_samples_is_a_clean_no_op(cuda, n_active):
    """Every centre ray in the zero region of the occupancy grid (models/renderer.py:137-140: the reference returns early and skips the
    optimizer step): the fused step must run through with S = 0 -- both backward forms, the tail kernel -- leave no NaN behind, produce zero
    gradients and, with fresh Adam moments, leave every parameter bit-identical.  NaN near / far (rays missing the unit sphere,
    models/dataset_loader.py:294-296) are part of the same batch."""
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer
    ds = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda)
    tr = FusedTrainer(ds, dict(DILIGENT_CONF, batch_size=128, end_iter=400, increase_bindwidth_every=10 ** 6), device=cuda)
    tr.iter_step = 101                       # not a multiple of the occupancy / bandwidth periods: the grid below stays as set
    tr.lr = 5e-4
    tr.model.n_active = n_active
    tr.grid._binary.zero_()
    before = tr.model.flat.clone()
    batch, jitter = tr.sample_batch_device(101)
    assert torch.isnan(batch["near"]).any()   # corner patches miss the unit sphere
    for _ in range(2):
        tr.train_step()
    lt = tr.loss_terms()
    assert lt["n_samples"] == 0 and lt["n_ends"] == 0 and lt["overflow"] == 0
    assert math.isfinite(lt["loss"]) and lt["eikonal"] == 0.0
    assert torch.equal(tr.model.flat, before) and (tr.model.grad == 0).all() and torch.isfinite(tr.model.net).all()
    assert (tr.buf.comp == 0).all() and (tr.buf.wsum == 0).all()
