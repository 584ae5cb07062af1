"""The reference's OWN models/renderer.py and models/fields.py -- the literal files, unmodified -- executed
(a) on the CPU over the oracle's operators: pins oracle/torch_ops.py's restatement of NeuSRenderer.render
    (models/renderer.py:63-276), SDFNetwork (models/fields.py:7-119) and SingleVarianceNetwork (:133-139);
(b) on the GPU over the product's drop-in modules supernormal_b200.nerfacc_api / tcnn_api (`-m gpu`): the
    drop-in claim itself -- the reference's model code runs unchanged on this library, and its outputs equal those of
    the same code over the UNMODIFIED reference nerfacc kernels (oracle/_ref) and those of the fused training step.
The files come from /root/reference where it exists, else from the byte-compiled oracle/_ref/models/*.pyc
(oracle/build_ref.py:build_models) -- the GPU box has no /root/reference."""
import math

import numpy as np
import pytest
import torch

from oracle import ref_models as R
from oracle import torch_ops as T
from supernormal_b200.synthetic import DILIGENT_CONF, SyntheticDataset, SyntheticScene

pytestmark = pytest.mark.skipif(not R.available(), reason="reference models neither at /root/reference nor byte-compiled in oracle/_ref/models")

SDF_KW = dict(d_out=1, d_in=3, d_hidden=64, n_layers=1, skip_in=[-1], bias=0.6, geometric_init=True, weight_norm=True,
              input_concat=True)   # config/diligent.conf:54-65


def _perturb(sdf_ref, seed=3):
    """geometric init zeroes the feature columns of lin0 (models/fields.py:57): give them weight so the encoding matters"""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        v = sdf_ref.lin0.weight_v
        v[:, 3:] = torch.randn(v[:, 3:].shape, generator=g).to(v) * 0.05
        p = sdf_ref.encoding.params
        p.copy_(((torch.rand(p.shape, generator=g) * 2 - 1) * 2e-2).to(p))


def _copy_into_oracle(sdf_ref, var_ref, sdf_o, var_o):
    sd = {k.replace("encoding.params", "encoding_params"): v.detach().cpu().clone() for k, v in sdf_ref.state_dict().items()}
    sdf_o.load_state_dict(sd)       # the oracle keeps the table as a plain parameter `encoding_params`
    with torch.no_grad():
        var_o.variance.copy_(var_ref.variance.detach().cpu())


def test_literal_reference_models_on_cpu_equal_the_oracle_restatement():
    enc_cfg = dict(DILIGENT_CONF["encoding"], n_levels=6, log2_hashmap_size=14)   # small table: seconds on the CPU
    rend_mod, fields_mod = R.load(R.cpu_nerfacc(), R.cpu_tcnn())
    torch.manual_seed(0)
    sdf_ref = fields_mod.SDFNetwork(**SDF_KW, encoding_config=enc_cfg)
    var_ref = fields_mod.SingleVarianceNetwork(init_val=0.5)
    _perturb(sdf_ref)
    sdf_ref.bindwidth = 4
    # state-dict keys the checkpoints carry (exp_runner.py:306-311)
    assert list(sdf_ref.state_dict().keys()) == ["encoding.params", "lin0.bias", "lin0.weight_g", "lin0.weight_v", "lin1.bias",
                                                  "lin1.weight_g", "lin1.weight_v"]
    ren_ref = rend_mod.NeuSRenderer(sdf_ref, var_ref, "dfd")

    sdf_o = T.SDFNetwork(enc_cfg, 64, 0.6)
    var_o = T.SingleVariance(0.5)
    _copy_into_oracle(sdf_ref, var_ref, sdf_o, var_o)
    sdf_o.bindwidth = 4
    ren_o = T.NeuSRenderer(sdf_o, var_o, "dfd")

    # SDFNetwork.forward / .gradient (models/fields.py:76-119)
    x = (torch.rand(257, 3) * 2 - 1) * 0.8
    assert torch.allclose(sdf_ref.sdf(x.clone()), sdf_o.sdf(x.clone()), atol=1e-6, rtol=0)
    g_ref = sdf_ref.gradient(x.clone())[:, 0]
    g_o = sdf_o.gradient(x.clone()).reshape(-1, 3)
    assert torch.allclose(g_ref, g_o, atol=1e-4, rtol=1e-4)   # fp16 output rounding (tcnn returns half) enters the literal path only

    ds = SyntheticDataset(SyntheticScene(n_views=4, H=64, W=80, exclude_views=(0,)), device="cpu")
    o, d, pn, vinv, nrm, msk = ds.gen_random_patches(48, 3, 3, np_rng=np.random.RandomState(1))
    near, far = ds.near_far_from_sphere(o[:, 1, 1], d[:, 1, 1])
    # occupancy grid through the literal renderer's own occ_eval_fn (models/renderer.py:56-60) and nerfacc's every_n_step
    torch.manual_seed(5)
    ren_ref.occupancy_grid.every_n_step(step=0, occ_eval_fn=ren_ref.occ_eval_fn, occ_thre=0.1, n=8)
    torch.manual_seed(5)
    ren_o.occupancy_grid.every_n_step(0, ren_o.occ_eval_fn, occ_thre=0.1, n=8)
    assert torch.equal(ren_ref.occupancy_grid.binary, ren_o.occupancy_grid.binary) and ren_o.occupancy_grid.binary.any()

    for method in ("dfd", "ad"):
        ren_ref.gradient_method = method
        ren_ref.sampling_step_size = ren_o.sampling_step_size = 0.02
        torch.manual_seed(11)
        out_ref = ren_ref.render(o, d, pn, near, far, vinv)
        torch.manual_seed(11)
        jitter = torch.rand_like(near)          # the draw NA/ray_marching.py:158 makes inside ray_marching
        out_o = ren_o.render(o, d, pn, near, far, vinv, jitter=jitter, gradient_method=method)
        S = out_o["n_samples"]
        assert S > 100 and out_ref["samples_per_ray"] == S / 48
        for k, tol in (("weight_sum", 1e-6), ("gradients", 2e-4), ("comp_normal", 2e-4), ("s_val", 1e-7)):
            a, b = out_ref[k].detach(), out_o[k].detach()
            if k == "s_val":
                a, b = a.reshape(()), b.reshape(())
            assert a.shape == b.shape, k
            assert (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item()), (method, k, (a - b).abs().max().item())
        # and the loss gradients that flow back through the literal code equal the restatement's
        loss_ref, _ = T.losses(out_ref, nrm, msk)
        loss_o, _ = T.losses(out_o, nrm, msk)
        assert abs(loss_ref.item() - loss_o.item()) <= 1e-5 * max(1.0, abs(loss_o.item()))
        for net in (sdf_ref, sdf_o, var_ref, var_o):
            net.zero_grad()
        loss_ref.backward()
        loss_o.backward()
        for (n1, p1), (n2, p2) in zip(sdf_ref.named_parameters(), sdf_o.named_parameters()):
            assert n1.replace(".", "_") == n2.replace(".", "_")
            scale = max(p2.grad.abs().max().item(), 1e-12)
            assert (p1.grad - p2.grad).abs().max().item() <= 2e-3 * scale, (method, n1)
        assert abs(var_ref.variance.grad.item() - var_o.variance.grad.item()) <= 2e-3 * max(abs(var_o.variance.grad.item()), 1e-9)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture
def cuda_default():
    """exp_runner.py:637 sets torch.set_default_tensor_type('torch.cuda.FloatTensor'); the literal renderer relies on it
    (models/renderer.py:108 `torch.zeros([1, 3])`, models/fields.py:139 `torch.ones([len(x), 1])`)."""
    torch.set_default_device("cuda")
    yield torch.device("cuda")
    torch.set_default_device("cpu")


@pytest.mark.gpu
def test_literal_reference_models_run_on_the_dropin_modules(cuda_default):
    from oracle import cuda_path as cp
    from supernormal_b200 import nerfacc_api, tcnn_api
    from supernormal_b200.trainer import FusedTrainer
    dev = cuda_default
    rend_mod, fields_mod = R.load(nerfacc_api, tcnn_api)          # `import nerfacc` / `import tinycudann as tcnn` resolve to the product
    enc_cfg = dict(DILIGENT_CONF["encoding"])
    torch.manual_seed(0)
    sdf_ref = fields_mod.SDFNetwork(**SDF_KW, encoding_config=enc_cfg).to(dev)
    var_ref = fields_mod.SingleVarianceNetwork(init_val=0.5).to(dev)
    assert isinstance(sdf_ref.encoding, tcnn_api.Encoding) and sdf_ref.encoding.params.numel() == 11_872_000
    _perturb(sdf_ref)
    sdf_ref.bindwidth = 5
    ren_ref = rend_mod.NeuSRenderer(sdf_ref, var_ref, "dfd")
    assert isinstance(ren_ref.occupancy_grid, nerfacc_api.OccupancyGrid)

    ds = SyntheticDataset(SyntheticScene(n_views=6, H=128, W=160, exclude_views=(0,)), device=dev)
    n = 256
    o, d, pn, vinv, nrm, msk = ds.gen_random_patches(n, 3, 3, generator=torch.Generator(device=dev).manual_seed(2), np_rng=np.random.RandomState(2))
    near, far = ds.near_far_from_sphere(o[:, 1, 1], d[:, 1, 1])
    torch.manual_seed(5)
    ren_ref.occupancy_grid.every_n_step(step=0, occ_eval_fn=ren_ref.occ_eval_fn, occ_thre=0.1, n=8)
    assert ren_ref.occupancy_grid.binary.any()
    ren_ref.sampling_step_size = np.float64(0.01)        # a NumPy float64, like exp_runner.py:149 hands it over
    torch.manual_seed(11)
    out = ren_ref.render(o, d, pn, near, far, vinv)
    S = int(round(out["samples_per_ray"] * n))
    assert S > 1000 and out["comp_normal"].shape == (n, 3, 3, 3) and out["weight_sum"].shape == (n, 3, 3, 1)
    loss, _ = T.losses(out, nrm, msk)
    loss.backward()
    g_lit = {k: p.grad.clone() for k, p in sdf_ref.named_parameters()}
    assert all(torch.isfinite(g).all() for g in g_lit.values()) and g_lit["encoding.params"].abs().sum() > 0

    # (1) the same literal code over the UNMODIFIED reference nerfacc kernels (oracle/_ref), same weights, grid and jitter
    ref_na = cp.load_ref_nerfacc()
    if ref_na is not None:
        rend2, fields2 = R.load(ref_na, tcnn_api)
        torch.manual_seed(0)
        sdf2 = fields2.SDFNetwork(**SDF_KW, encoding_config=enc_cfg).to(dev)
        sdf2.load_state_dict(sdf_ref.state_dict())
        sdf2.bindwidth = 5
        ren2 = rend2.NeuSRenderer(sdf2, var_ref, "dfd")
        ren2.occupancy_grid._binary.copy_(ren_ref.occupancy_grid.binary) if hasattr(ren2.occupancy_grid, "_binary") else None
        ren2.occupancy_grid.binary = ren_ref.occupancy_grid.binary.clone() if not hasattr(ren2.occupancy_grid, "_binary") else ren2.occupancy_grid.binary
        ren2.sampling_step_size = np.float64(0.01)
        torch.manual_seed(11)
        out2 = ren2.render(o, d, pn, near, far, vinv)
        assert out2["samples_per_ray"] == out["samples_per_ray"]          # bit-exact marching + visibility: the same samples
        for k in ("weight_sum", "comp_normal", "gradients"):
            assert torch.allclose(out[k], out2[k], atol=1e-6, rtol=1e-5), k

    # (2) the fused training step's forward on the same weights / grid / batch / jitter: rendered normals within the north star's
    #     1e-3 rad mean angular error
    conf = dict(DILIGENT_CONF, batch_size=n)
    tr = FusedTrainer(ds, conf, device=dev)
    tr.model.load_reference_state_dict({"sdf_network_fine": sdf_ref.state_dict(), "variance_network_fine": var_ref.state_dict()})
    tr.model.n_active = 5
    tr.grid._binary.copy_(ren_ref.occupancy_grid.binary)
    torch.manual_seed(11)
    jitter = torch.rand_like(near)
    batch = dict(rays_o=o[:, 1, 1].contiguous(), rays_d=d.reshape(-1, 9, 3).contiguous(), plane_n=pn.contiguous(), near=near.contiguous(),
                 far=far.contiguous(), v_inv=vinv.reshape(-1, 9, 9).contiguous(), normal_gt=nrm.reshape(-1, 9, 3).contiguous(),
                 mask=msk.reshape(-1, 9).contiguous())
    tr.forward_backward(batch, 0.01, jitter, lean=False)
    assert int(tr.buf.totals[0].item()) == S
    comp_f = tr.buf.comp.view(n, 3, 3, 3)
    cn, cf = out["comp_normal"].detach(), comp_f
    m = (msk.reshape(n, 3, 3) > 0.5) & (cn.norm(dim=-1) > 0.1)
    cos = (torch.nn.functional.normalize(cn, dim=-1) * torch.nn.functional.normalize(cf, dim=-1)).sum(-1).clamp(-1, 1)
    assert torch.acos(cos)[m].mean().item() < 1e-3
    assert torch.allclose(tr.buf.wsum.view(n, 3, 3, 1), out["weight_sum"], atol=2e-4)
    # hash-table gradient of the fused backward vs autograd through the literal code (same loss)
    gt = tr.model.grad[tr.model.flat.numel() - tr.model.n_table:]
    ref_g = g_lit["encoding.params"]
    live = tr.model.offsets[5] * 2
    assert (gt[live:] == 0).all()
    err = (gt[:live] - ref_g[:live]).abs().max().item()
    assert err <= 5e-3 * ref_g[:live].abs().max().item(), err
