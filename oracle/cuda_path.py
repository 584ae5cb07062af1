"""The reference-SHAPED training step on CUDA.  TEST / MEASUREMENT INFRASTRUCTURE ONLY (see
oracle/__init__.py): never imported by the product.

The step is the same restatement of models/renderer.py:63-276 + models/fields.py:76-99 +
exp_runner.py:147-210 as oracle/torch_ops.py (ATen ops + torch.autograd + torch.optim.Adam, one
launch per ATen op like the reference), but on a CUDA device and written against the *public nerfacc /
tinycudann API* of a pluggable module pair:

  backend "reference":  nerfacc = `RefNerfacc` -- the nerfacc 0.3.5 Python layer restated over the
      UNMODIFIED reference CUDA kernels in oracle/_ref/nerfacc_ref_C.so (ray_marching, CUB
      transmittance, patch weights fwd/bwd; pack_info / accumulate / OccupancyGrid are ATen ops in
      the reference too).  tiny-cuda-nn is absent from this image, so the encoding is
      `supernormal_b200.tcnn_api.Encoding` behind the reference's calling convention
      (all levels computed, mask applied afterwards, fp32 master cast to fp16 every forward --
      models/fields.py:78-83).  This is the closest runnable stand-in for "the reference's own
      tcnn+nerfacc path on one B200" and what bench.py reports as `reference_cuda_path`.
  backend "dropin":     nerfacc = supernormal_b200.nerfacc_api, tcnn = supernormal_b200.tcnn_api --
      what a SuperNormal user gets by swapping the two imports, without the fused
      trainer.  tests/test_gpu_dropin.py checks it against the fused path and the reference kernels.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch
import torch.nn as nn

from oracle import torch_ops as T


# ----------------------------------------------------------------------------------------------
# nerfacc 0.3.5 Python layer over the reference's compiled kernels
# ----------------------------------------------------------------------------------------------
class RefNerfacc:
    """NA/ray_marching.py:14-222, NA/vol_rendering.py:269-335,533-576,680-748,990-1010, NA/pack.py:47-77
    with `_C` = oracle/_ref/nerfacc_ref_C.so (csrc/pybind.cu:162-206)."""

    def __init__(self, C):
        self._C = C
        outer = self

        class _W(torch.autograd.Function):
            @staticmethod
            def forward(ctx, packed_info, alphas):
                packed_info, alphas = packed_info.contiguous(), alphas.contiguous()
                w = outer._C.weight_from_alpha_patch_based_forward_naive(packed_info, alphas)
                if ctx.needs_input_grad[1]:
                    ctx.save_for_backward(packed_info, alphas, w)
                return w

            @staticmethod
            def backward(ctx, gw):
                packed_info, alphas, w = ctx.saved_tensors
                return None, outer._C.weight_from_alpha_patch_based_backward_naive(w, gw.contiguous(), packed_info, alphas)

        self._W = _W

    @staticmethod
    def pack_info(ray_indices, n_rays=None):
        if n_rays is None:
            n_rays = int(ray_indices.max()) + 1
        src = torch.ones_like(ray_indices, dtype=torch.int)
        num = torch.zeros((n_rays,), device=ray_indices.device, dtype=torch.int)
        num.scatter_add_(0, ray_indices.long(), src)
        cum = num.cumsum(dim=0, dtype=torch.int)
        return torch.stack([cum - num, num], dim=-1)

    def render_visibility(self, alphas, *, ray_indices=None, packed_info=None, early_stop_eps=1e-4, alpha_thre=0.0, n_rays=None):
        T_ = self._C.transmittance_from_alpha_forward_cub(ray_indices.contiguous(), alphas.contiguous())
        vis = T_ >= early_stop_eps
        if alpha_thre > 0:
            vis = vis & (alphas >= alpha_thre)
        return vis.squeeze(-1)

    @torch.no_grad()
    def ray_marching(self, rays_o, rays_d, t_min=None, t_max=None, grid=None, alpha_fn: Optional[Callable] = None,
                     early_stop_eps=1e-4, alpha_thre=0.0, render_step_size=1e-3, stratified=False, cone_angle=0.0):
        if stratified:
            t_min = t_min + torch.rand_like(t_min) * render_step_size
        packed_info, ray_indices, t0, t1 = self._C.ray_marching(
            rays_o.contiguous(), rays_d.contiguous(), t_min.contiguous(), t_max.contiguous(), grid.roi_aabb.contiguous(),
            grid.binary.contiguous(), self._C.ContractionType.AABB, float(render_step_size), float(cone_angle))
        if alpha_fn is not None:
            alphas = alpha_fn(t0, t1, ray_indices)
            m = self.render_visibility(alphas, ray_indices=ray_indices, packed_info=packed_info, early_stop_eps=early_stop_eps,
                                       alpha_thre=alpha_thre, n_rays=rays_o.shape[0])
            ray_indices, t0, t1 = ray_indices[m], t0[m], t1[m]
        return ray_indices, t0, t1

    def render_weight_from_alpha_patch_based(self, alphas, ray_indices, *, n_rays=None):
        return self._W.apply(self.pack_info(ray_indices, n_rays), alphas)

    @staticmethod
    def accumulate_along_rays_patch_based(weights, ray_indices, values=None, n_patches=None):
        src = weights * values if values is not None else weights
        if ray_indices.numel() == 0:
            return torch.zeros((n_patches, src.shape[1], src.shape[-1]), device=weights.device)
        if n_patches is None:
            n_patches = int(ray_indices.max()) + 1
        index = ray_indices[:, None, None].expand(-1, src.shape[1], src.shape[-1])
        out = torch.zeros((n_patches, src.shape[1], src.shape[-1]), device=src.device, dtype=src.dtype)
        out.scatter_add_(0, index, src)
        return out

    def OccupancyGrid(self, roi_aabb, resolution=128, device="cuda"):
        return T.OccupancyGrid(roi_aabb, resolution, device=device)   # NA/grid.py is ATen ops only


def load_ref_nerfacc() -> Optional[RefNerfacc]:
    from oracle import build_ref
    C = build_ref.load()
    return None if C is None else RefNerfacc(C)


# ----------------------------------------------------------------------------------------------
# models/fields.py on a tcnn-shaped encoding module
# ----------------------------------------------------------------------------------------------
class SDFNetworkTcnn(nn.Module):
    """models/fields.py:7-99 (shipped configuration) with `tcnn.Encoding` from the given module.
    `reference_convention=True` reproduces what the reference pays for with tiny-cuda-nn: every level is
    evaluated and the inactive ones are masked afterwards (models/fields.py:81-83), and the fp32 master
    table is cast to fp16 on every forward (tcnn's torch binding).  False uses the drop-in module's extras
    (`n_active_levels`, version-cached fp16 table)."""

    def __init__(self, tcnn, encoding_config: dict, d_hidden=64, bias=0.6, seed=1337, reference_convention=False):
        super().__init__()
        self.encoding = tcnn.Encoding(3, encoding_config, seed=seed)
        self.n_levels = int(encoding_config["n_levels"])
        self.enc_dim = self.encoding.n_output_dims
        d0 = 3 + self.enc_dim
        g = torch.Generator().manual_seed(seed)
        _ = torch.rand(self.encoding.params.numel(), generator=g)       # same stream position as oracle.torch_ops.SDFNetwork
        lin0, lin1 = nn.Linear(d0, d_hidden), nn.Linear(d_hidden, 1)
        with torch.no_grad():
            lin0.bias.zero_()
            lin0.weight[:, 3:].zero_()
            lin0.weight[:, :3].normal_(0.0, math.sqrt(2) / math.sqrt(d_hidden), generator=g)
            lin1.weight.normal_(math.sqrt(math.pi) / math.sqrt(d_hidden), 1e-4, generator=g)
            lin1.bias.fill_(-bias)
        self.lin0 = nn.utils.weight_norm(lin0)
        self.lin1 = nn.utils.weight_norm(lin1)
        self.bindwidth = 0
        self.reference_convention = reference_convention
        self.register_buffer("mask", torch.zeros(self.enc_dim), persistent=False)

    def increase_bandwidth(self):
        self.bindwidth += 1

    def forward(self, x):
        if self.reference_convention:
            self.encoding._cache_key = None          # tcnn casts the fp32 master to fp16 on every call
            self.encoding.n_active_levels = None
            enc = self.encoding(x).to(torch.float32)
            mask = torch.zeros(self.enc_dim, device=x.device)
            mask[: self.bindwidth * 2] = 1.0
            enc = enc * mask
        else:
            self.encoding.n_active_levels = self.bindwidth
            enc = self.encoding(x).to(torch.float32)
        h = T.softplus100(self.lin0(torch.cat([x, enc], 1)))
        return self.lin1(h)

    def sdf(self, x):
        return self.forward(x)[:, :1]

    @torch.enable_grad()
    def gradient(self, x):
        x.requires_grad_(True)
        y = self.sdf(x)
        (g,) = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True)
        return g.unsqueeze(1)


# ----------------------------------------------------------------------------------------------
# renderer + loop
# ----------------------------------------------------------------------------------------------
class CudaRenderer(T.NeuSRenderer):
    """torch_ops.NeuSRenderer with marching / weights / accumulation from a nerfacc-shaped module."""

    def __init__(self, nerfacc, sdf_network, deviation, gradient_method="dfd", device="cuda"):
        super().__init__(sdf_network, deviation, gradient_method, ops=nerfacc, device=device)
        self.nerfacc = nerfacc
        g = nerfacc.OccupancyGrid(self.scene_aabb, 128)
        self.occupancy_grid = g.to(device) if isinstance(g, nn.Module) else nerfacc.OccupancyGrid(self.scene_aabb, 128, device=device)

    def march(self, o_c, d_c, near, far, jitter):
        """models/renderer.py:124-135 (stratified jitter drawn inside ray_marching, like the reference)."""
        return self.nerfacc.ray_marching(o_c, d_c, t_min=near, t_max=far, grid=self.occupancy_grid,
                                         render_step_size=np.float32(self.sampling_step_size), stratified=True, cone_angle=0.0,
                                         early_stop_eps=1e-8, alpha_fn=self.centre_alpha_fn(o_c, d_c))


class CudaTrainer(T.Trainer):
    """exp_runner.py:83-210 on CUDA; `backend` in {"reference", "dropin"}."""

    def __init__(self, dataset, conf: dict, backend: str = "dropin", seed=0, device="cuda"):
        from supernormal_b200 import tcnn_api
        if backend == "reference":
            nerfacc = load_ref_nerfacc()
            if nerfacc is None:
                raise RuntimeError("oracle/_ref/nerfacc_ref_C.so is not built (oracle/build_ref.py needs /root/reference)")
        elif backend == "dropin":
            from supernormal_b200 import nerfacc_api as nerfacc
        else:
            raise ValueError(backend)
        self.backend = backend
        self.ds, self.conf = dataset, conf
        torch.manual_seed(seed)
        self.np_rng = np.random.RandomState(seed)
        self.sdf = SDFNetworkTcnn(tcnn_api, conf["encoding"], conf["sdf_network"]["d_hidden"], conf["sdf_network"]["bias"],
                                  reference_convention=(backend == "reference")).to(device)
        self.dev = T.SingleVariance(conf["variance_init"]).to(device)
        self.renderer = CudaRenderer(nerfacc, self.sdf, self.dev, conf["gradient_method"], device=device)
        self.opt = torch.optim.Adam(list(self.sdf.parameters()) + list(self.dev.parameters()), lr=conf["learning_rate"])
        rm = conf["ray_marching"]
        self.slop = (math.log10(rm["start_step_size"]) - math.log10(rm["end_step_size"])) / conf["end_iter"]
        self.iter_step = 0
        self._set_lr()      # exp_runner.py:125: update_learning_rate() before the first iteration -> step 0 runs at lr = 0


def time_steps(tr: CudaTrainer, steps: int, warmup: int):
    """-> (ms per step via CUDA events, samples-per-ray of the last step)."""
    for _ in range(warmup):
        tr.step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out = None
    for _ in range(steps):
        _, out = tr.step()
    e1.record()
    torch.cuda.synchronize()
    spr = (out["n_samples"] / tr.conf["batch_size"]) if out else 0.0
    return e0.elapsed_time(e1) / steps, spr
