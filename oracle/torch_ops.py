"""PyTorch-CPU expression of the patch-based NeuS training step.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py): parity checker for the CUDA path and the reported CPU baseline.

Each function restates the reference file:line it cites (paths relative to /root/reference;
NA = third_parties/nerfacc-0.3.5/nerfacc-0.3.5/nerfacc).  Kernels whose rounding matters
(ray marching, rendering weights) go through the bit-faithful C restatement in
oracle/c/oracle.c; everything the reference itself computes with ATen ops is computed with
the same ATen ops here.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import oracle as _o

_M32 = 0xFFFFFFFF


# ----------------------------------------------------------------------------------------
# nerfacc Python layer
# ----------------------------------------------------------------------------------------
def pack_info(ray_indices: torch.Tensor, n_rays: Optional[int] = None) -> torch.Tensor:
    """NA/pack.py:47-77."""
    assert ray_indices.dim() == 1
    if n_rays is None:
        n_rays = int(ray_indices.max()) + 1 if ray_indices.numel() else 0
    num = torch.zeros(n_rays, dtype=torch.int32)
    num.scatter_add_(0, ray_indices.long(), torch.ones_like(ray_indices, dtype=torch.int32))
    cum = num.cumsum(0, dtype=torch.int32)
    return torch.stack([cum - num, num], dim=-1)


def render_transmittance_from_alpha(alphas, *, ray_indices=None, packed_info=None, n_rays=None):
    """Serial exclusive product (CS/render_transmittance.cu:85-112)."""
    if packed_info is None:
        packed_info = pack_info(ray_indices, n_rays)
    T = _o.transmittance_from_alpha(packed_info.numpy(), alphas.detach().numpy())
    return torch.from_numpy(T)


def render_visibility(alphas, *, ray_indices=None, packed_info=None, n_rays=None,
                      early_stop_eps=1e-4, alpha_thre=0.0):
    """NA/vol_rendering.py:680-748."""
    T = render_transmittance_from_alpha(alphas, ray_indices=ray_indices, packed_info=packed_info, n_rays=n_rays)
    vis = T >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis.squeeze(-1)


class _WeightFromAlphaPatch(torch.autograd.Function):
    """NA/vol_rendering.py:990-1010 over CS/render_weight.cu:87-118,298-340."""

    @staticmethod
    def forward(ctx, packed_info, alphas):
        w = torch.from_numpy(_o.weight_from_alpha_patch_fwd(packed_info.numpy(), alphas.detach().numpy()))
        ctx.save_for_backward(packed_info, alphas.detach(), w)
        return w

    @staticmethod
    def backward(ctx, gw):
        packed_info, alphas, w = ctx.saved_tensors
        g = _o.weight_from_alpha_patch_bwd(packed_info.numpy(), alphas.numpy(), w.numpy(), gw.contiguous().numpy())
        return None, torch.from_numpy(g)


def render_weight_from_alpha_patch_based(alphas, ray_indices, *, n_rays=None):
    """NA/vol_rendering.py:533-576. alphas [S,P,1]."""
    return _WeightFromAlphaPatch.apply(pack_info(ray_indices, n_rays), alphas.contiguous())


def render_weight_from_alpha(alphas, *, ray_indices=None, packed_info=None, n_rays=None):
    """NA/vol_rendering.py:624-677 (naive route == patch route with P=1)."""
    if packed_info is None:
        packed_info = pack_info(ray_indices, n_rays)
    return _WeightFromAlphaPatch.apply(packed_info, alphas.contiguous())


def accumulate_along_rays_patch_based(weights, ray_indices, values=None, n_patches=None):
    """NA/vol_rendering.py:269-335."""
    assert ray_indices.dim() == 1 and weights.dim() == 3
    src = weights * values if values is not None else weights
    if ray_indices.numel() == 0:
        return torch.zeros((n_patches, src.shape[1], src.shape[-1]))
    if n_patches is None:
        n_patches = int(ray_indices.max()) + 1
    out = torch.zeros((n_patches, src.shape[1], src.shape[-1]), dtype=src.dtype)
    return out.index_add(0, ray_indices.long(), src)


def accumulate_along_rays(weights, ray_indices, values=None, n_rays=None):
    """NA/vol_rendering.py:132-198."""
    assert ray_indices.dim() == 1 and weights.dim() == 2
    src = weights * values if values is not None else weights
    if ray_indices.numel() == 0:
        return torch.zeros((n_rays, src.shape[-1]))
    if n_rays is None:
        n_rays = int(ray_indices.max()) + 1
    return torch.zeros((n_rays, src.shape[-1]), dtype=src.dtype).index_add(0, ray_indices.long(), src)


def ray_marching(rays_o, rays_d, t_min, t_max, grid_roi, grid_binary, render_step_size,
                 cone_angle=0.0, alpha_fn: Optional[Callable] = None, early_stop_eps=1e-4,
                 alpha_thre=0.0, jitter: Optional[torch.Tensor] = None):
    """NA/ray_marching.py:145-222.  `jitter` (uniform [0,1) per ray) replaces the
    torch.rand_like of :158 so tests can inject it; None = not stratified."""
    if jitter is not None:
        t_min = t_min + jitter * render_step_size
    packed, ridx, t0, t1 = _o.ray_marching(rays_o.numpy(), rays_d.numpy(), t_min.numpy(), t_max.numpy(),
                                            grid_roi.numpy(), grid_binary.numpy(), float(render_step_size), cone_angle)
    packed, ridx, t0, t1 = map(torch.from_numpy, (packed, ridx, t0, t1))
    if alpha_fn is not None and ridx.numel() > 0:
        alphas = alpha_fn(t0, t1, ridx)
        m = render_visibility(alphas, packed_info=packed, early_stop_eps=early_stop_eps, alpha_thre=alpha_thre)
        ridx, t0, t1 = ridx[m], t0[m], t1[m]
    return ridx, t0, t1


class OccupancyGrid:
    """NA/grid.py:113-294 (AABB contraction only: contract_inv = x*(max-min)+min,
    CS/include/helpers_contraction.h:23-28)."""

    def __init__(self, roi_aabb, resolution=128, device="cpu"):
        self.device = torch.device(device)
        self.roi_aabb = torch.as_tensor(roi_aabb, dtype=torch.float32).to(self.device)
        self.resolution = torch.tensor([resolution] * 3, dtype=torch.int32, device=self.device)
        self.num_cells = resolution ** 3
        self.binary = torch.zeros([resolution] * 3, dtype=torch.bool, device=self.device)
        self.occs = torch.zeros(self.num_cells, device=self.device)
        r = torch.arange(resolution, device=self.device)
        self.grid_coords = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), -1).reshape(-1, 3)

    def update(self, step, occ_eval_fn, occ_thre=0.01, ema_decay=0.95, warmup_steps=256,
               rand: Optional[torch.Tensor] = None, indices: Optional[torch.Tensor] = None):
        """NA/grid.py:197-239. `rand` [n,3] / `indices` injectable for parity tests."""
        if indices is None:
            if step < warmup_steps:
                indices = torch.arange(self.num_cells, device=self.device)
            else:
                N = self.num_cells // 4
                uni = torch.randint(self.num_cells, (N,), device=self.device)
                occ_idx = torch.nonzero(self.binary.flatten())[:, 0]
                if N < len(occ_idx):
                    occ_idx = occ_idx[torch.randint(len(occ_idx), (N,), device=self.device)]
                indices = torch.cat([uni, occ_idx])
        coords = self.grid_coords[indices]
        if rand is None:
            rand = torch.rand(coords.shape, device=self.device)
        x = (coords + rand) / self.resolution
        x = x * (self.roi_aabb[3:] - self.roi_aabb[:3]) + self.roi_aabb[:3]
        occ = occ_eval_fn(x).squeeze(-1)
        self.occs[indices] = torch.maximum(self.occs[indices] * ema_decay, occ)
        self.binary = (self.occs > torch.clamp(self.occs.mean(), max=occ_thre)).view(self.binary.shape)

    def every_n_step(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16, **kw):
        if step % n == 0:
            self.update(step, occ_eval_fn, occ_thre, ema_decay, warmup_steps, **kw)


# ----------------------------------------------------------------------------------------
# tiny-cuda-nn HashGrid (parity unpinned) -- differentiable to 2nd order in x
# ----------------------------------------------------------------------------------------
def _ste_half(t: torch.Tensor) -> torch.Tensor:
    """Round to fp16 in the forward value, identity in the gradient."""
    return t + (t.detach().half().to(t.dtype) - t.detach())


def _grid_index(size: int, res: int, p):
    """tcnn grid_index (common_device.h) in wrapped-uint32 arithmetic on int64 tensors."""
    stride, index, dim = 1, torch.zeros_like(p[0]), 0
    while dim < 3 and stride <= size:
        index = index + p[dim] * stride
        stride = (stride * res) & _M32
        dim += 1
    if size < stride:
        index = p[0] ^ ((p[1] * 2654435761) & _M32) ^ ((p[2] * 805459861) & _M32)
    return (index & _M32) % size


def hashgrid_encode(x: torch.Tensor, params: torch.Tensor, spec: "_o.HashGridSpec",
                    n_active: Optional[int] = None, fp16: bool = True) -> torch.Tensor:
    """tcnn kernel_grid with ATen ops.  fp16=True rounds the table
    and the output to fp16 (straight-through) like the reference's fp16 params/outputs; the
    fp16 *accumulation* order is only emulated by the C oracle (oracle.hashgrid_fwd)."""
    n_active = spec.n_levels if n_active is None else n_active
    tbl = params.view(-1, 2)
    if fp16:
        tbl = _ste_half(tbl)
    tbl = tbl.to(x.dtype)
    outs = []
    for l in range(spec.n_levels):
        if l >= n_active:
            outs.append(torch.zeros(x.shape[0], 2, dtype=x.dtype))
            continue
        scale, res = float(spec.scales[l]), int(spec.resolutions[l])
        off, size = int(spec.offsets[l]), int(spec.offsets[l + 1] - spec.offsets[l])
        pos = x * scale + 0.5
        fl = torch.floor(pos.detach())
        w = pos - fl
        cell = fl.long() & _M32
        feat = 0
        for c in range(8):
            wt, pl = 1.0, []
            for d in range(3):
                if c & (1 << d):
                    wt = wt * w[:, d]
                    pl.append((cell[:, d] + 1) & _M32)
                else:
                    wt = wt * (1 - w[:, d])
                    pl.append(cell[:, d])
            feat = feat + wt[:, None] * tbl[off + _grid_index(size, res, pl)]
        outs.append(feat)
    out = torch.cat(outs, 1)
    if not fp16:
        return out
    if x.dtype == torch.float32:
        # forward VALUE from the fp16-faithful C restatement (fp16 weights, fp16 fma chain), gradient of the
        # fp32 expression above (straight-through, like the fp16 rounding itself)
        exact = _o.hashgrid_fwd(spec, x.detach().numpy(), params.detach().numpy().astype(np.float16), n_active=n_active)
        return out + (torch.from_numpy(exact.astype(np.float32)) - out.detach())
    return _ste_half(out)


def softplus100(x):
    """nn.Softplus(beta=100), threshold 20 (models/fields.py:70)."""
    return F.softplus(x, beta=100.0, threshold=20.0)


class SDFNetwork(nn.Module):
    """models/fields.py:7-119 for the shipped configuration (n_layers=1, no skip, weight-norm,
    geometric init, input_concat) with the tcnn encoding replaced by `hashgrid_encode`.
    State-dict keys equal the reference's: encoding_params <-> encoding.params, lin{0,1}.*"""

    def __init__(self, encoding_config: dict, d_hidden=64, bias=0.6, seed=1337, fp16=True, dtype=torch.float32):
        super().__init__()
        self.spec = _o.hashgrid_spec(**{k: v for k, v in encoding_config.items() if k != "otype"})
        g = torch.Generator().manual_seed(seed)
        self.encoding_params = nn.Parameter((torch.rand(self.spec.n_params, generator=g) * 2 - 1) * 1e-4)
        self.enc_dim = self.spec.n_output_dims
        d0 = 3 + self.enc_dim
        lin0, lin1 = nn.Linear(d0, d_hidden), nn.Linear(d_hidden, 1)
        with torch.no_grad():  # geometric init, models/fields.py:47-65
            lin0.bias.zero_()
            lin0.weight[:, 3:].zero_()
            lin0.weight[:, :3].normal_(0.0, math.sqrt(2) / math.sqrt(d_hidden), generator=g)
            lin1.weight.normal_(math.sqrt(math.pi) / math.sqrt(d_hidden), 1e-4, generator=g)
            lin1.bias.fill_(-bias)
        self.lin0 = nn.utils.weight_norm(lin0)
        self.lin1 = nn.utils.weight_norm(lin1)
        self.bindwidth = 0
        self.fp16 = fp16

    def increase_bandwidth(self):
        self.bindwidth += 1

    def forward(self, x):
        enc = hashgrid_encode(x, self.encoding_params, self.spec, n_active=self.bindwidth, fp16=self.fp16)
        h = softplus100(self.lin0(torch.cat([x, enc.to(x.dtype)], 1)))
        return self.lin1(h)

    def sdf(self, x):
        return self.forward(x)[:, :1]

    @torch.enable_grad()
    def gradient(self, x):
        """models/fields.py:107-119."""
        x.requires_grad_(True)
        y = self.sdf(x)
        (g,) = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True)
        return g.unsqueeze(1)


class SingleVariance(nn.Module):
    """models/fields.py:133-139."""

    def __init__(self, init_val=0.5):
        super().__init__()
        self.variance = nn.Parameter(torch.tensor(float(init_val)))

    def inv_s(self):
        return torch.exp(self.variance * 10.0).clip(1e-6, 1e6)


# ----------------------------------------------------------------------------------------
# models/renderer.py
# ----------------------------------------------------------------------------------------
def neus_alpha(sdf0, sdf1, inv_s):
    """models/renderer.py:173-179."""
    c = torch.sigmoid(sdf0 * inv_s)
    n = torch.sigmoid(sdf1 * inv_s)
    return ((c - n + 1e-5) / (c + 1e-5)).clip(0.0, 1.0)


def _next_start_or_own_end(v_start, v_end_diff, diff_mask):
    """models/renderer.py:164-169: value at the interval end = next interval's start value
    unless the two intervals are not contiguous."""
    nxt = torch.cat([v_start[1:], v_start[-1:]], 0).clone()
    nxt[diff_mask] = v_end_diff
    return nxt


def _diff_mask(t0, t1):
    return ((t1 - torch.cat([t0[1:], t0[-1:]], 0)) != 0)[..., 0]


class NeuSRenderer:
    """models/renderer.py:37-276 (patch-based `render`), CPU, our oracle operators."""

    def __init__(self, sdf_network: SDFNetwork, deviation: SingleVariance, gradient_method="dfd", ops=None, device="cpu"):
        """`ops` = namespace providing render_weight_from_alpha_patch_based / accumulate_along_rays_patch_based
        (default: this module's CPU oracle operators; oracle/cuda_path.py plugs in a nerfacc-shaped CUDA module)."""
        import sys
        self.ops = ops if ops is not None else sys.modules[__name__]
        self.sdf_network, self.deviation_network = sdf_network, deviation
        self.scene_aabb = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], device=device)
        self.occupancy_grid = OccupancyGrid(self.scene_aabb, 128, device=device)
        self.sampling_step_size = 0.01
        self.gradient_method = gradient_method

    def occ_eval_fn(self, x):
        """models/renderer.py:56-60."""
        with torch.no_grad():
            return torch.sigmoid(-self.sdf_network(x)[..., :1] * 80)

    def centre_alpha_fn(self, o_c, d_c):
        """models/renderer.py:80-122."""
        def fn(t0, t1, ridx):
            with torch.no_grad():
                o, d = o_c[ridx.long()], d_c[ridx.long()]
                ps, pe = o + d * t0, o + d * t1
                dm = _diff_mask(t0, t1)
                sdf = self.sdf_network(torch.cat([ps, pe[dm].reshape(-1, 3)], 0))
                s0 = sdf[: ps.shape[0]]
                s1 = _next_start_or_own_end(s0, sdf[ps.shape[0]:], dm)
                return neus_alpha(s0, s1, self.deviation_network.inv_s()).reshape(-1, 1)
        return fn

    def march(self, o_c, d_c, near, far, jitter):
        return ray_marching(o_c, d_c, near, far, self.scene_aabb, self.occupancy_grid.binary,
                            np.float32(self.sampling_step_size), 0.0, self.centre_alpha_fn(o_c, d_c),
                            early_stop_eps=1e-8, jitter=jitter)

    def render(self, rays_o, rays_d, plane_n, near, far, V_inv, jitter=None, samples=None, gradient_method=None):
        """models/renderer.py:63-276.  `samples` = (patch_idx, t0, t1) skips marching."""
        Np, pH, pW = rays_o.shape[:3]
        o_c, d_c = rays_o[:, pH // 2, pW // 2], rays_d[:, pH // 2, pW // 2]
        with torch.no_grad():
            pidx, t0c, t1c = samples if samples is not None else self.march(o_c, d_c, near, far, jitter)
            pidx = pidx.long()
            S = pidx.shape[0]
            if S == 0:
                return {"comp_normal": torch.zeros(Np, pH, pW, 3, device=rays_o.device), "gradients": None, "n_samples": 0}
            # plane fan-out, models/renderer.py:146-149
            num = (d_c * plane_n).sum(-1, keepdim=True)[pidx][:, None, None, :]
            den = (rays_d * plane_n[:, None, None, :]).sum(-1, keepdim=True)[pidx]
            t0 = t0c[:, None, None, :] * num / den
            t1 = t1c[:, None, None, :] * num / den
            dm = _diff_mask(t0c, t1c)
            p0 = rays_o[pidx] + rays_d[pidx] * t0
            p1 = rays_o[pidx] + rays_d[pidx] * t1
            pos_all = torch.cat([p0, p1[dm]], 0)
        sdf_all = self.sdf_network(pos_all.reshape(-1, 3)).reshape(*pos_all.shape[:-1], 1)
        if sdf_all.requires_grad:
            sdf_all.retain_grad()  # tests compare d loss / d sdf with the fused kernels' seeds
        s0 = sdf_all[:S]
        s1 = _next_start_or_own_end(s0, sdf_all[S:], dm)
        inv_s = self.deviation_network.inv_s()
        alpha = neus_alpha(s0, s1, inv_s)
        w = self.ops.render_weight_from_alpha_patch_based(alpha.reshape(S, pH * pW, 1), pidx)
        method = gradient_method or self.gradient_method
        if method == "dfd":  # models/renderer.py:187-223
            with torch.no_grad():
                dist_x = (p0[:, :, 1:] - p0[:, :, :-1]).norm(dim=-1, keepdim=True)
                dist_y = (p0[:, 1:] - p0[:, :-1]).norm(dim=-1, keepdim=True)
            df_dt = (s1 - s0) / (t1 - t0)
            dx_c = (s0[:, :, 2:] - s0[:, :, :-2]) / (dist_x[:, :, :-1] + dist_x[:, :, 1:])
            dy_c = (s0[:, 2:] - s0[:, :-2]) / (dist_y[:, 1:] + dist_y[:, :-1])
            dx_l = (s0[:, :, 1:2] - s0[:, :, 0:1]) / dist_x[:, :, 0:1]
            dx_r = (s0[:, :, -1:] - s0[:, :, -2:-1]) / dist_x[:, :, -1:]
            dy_t = (s0[:, 1:2] - s0[:, 0:1]) / dist_y[:, 0:1]
            dy_b = (s0[:, -1:] - s0[:, -2:-1]) / dist_y[:, -1:]
            proj = torch.cat([df_dt, torch.cat([dx_l, dx_c, dx_r], 2), torch.cat([dy_t, dy_c, dy_b], 1)], -1)
            grads = (V_inv[pidx] @ proj[..., None])[..., 0]
        elif method == "ad":
            grads = self.sdf_network.gradient(p0.reshape(-1, 3)).reshape(S, pH, pW, 3)
        else:
            raise ValueError(method)
        wsum = self.ops.accumulate_along_rays_patch_based(w, pidx, n_patches=Np).reshape(Np, pH, pW, 1)
        comp = self.ops.accumulate_along_rays_patch_based(w, pidx, values=grads.reshape(S, pH * pW, 3), n_patches=Np)
        return {"s_val": 1 / inv_s, "weight_sum": wsum, "gradients": grads, "comp_normal": comp.reshape(Np, pH, pW, 3),
                "n_samples": S, "samples": (pidx, t0c, t1c), "sdf_all": sdf_all, "diff_mask": dm, "sdf_start": s0, "sdf_end": s1, "alpha": alpha, "weights": w}


def losses(out, true_normal, mask, normal_weight=1.0, mask_weight=1.0, eikonal_weight=1.0):
    """exp_runner.py:169-203 (l2)."""
    mask = (mask > 0.5).float() if mask_weight > 0 else torch.ones_like(mask)
    mask_sum = mask.sum() + 1e-5
    err = (out["comp_normal"] - true_normal) * mask
    normal_loss = F.mse_loss(err, torch.zeros_like(err), reduction="sum") / mask_sum
    gn = torch.linalg.norm(out["gradients"], ord=2, dim=-1)
    eik = F.mse_loss(gn, torch.ones_like(gn), reduction="mean")
    mloss = F.binary_cross_entropy(out["weight_sum"].clip(1e-5, 1.0 - 1e-5), mask)
    total = normal_weight * normal_loss + mask_weight * mloss + eikonal_weight * eik
    return total, {"normal": normal_loss, "mask": mloss, "eikonal": eik}


def lr_factor(iter_step, warm_up_end, end_iter, alpha):
    """exp_runner.py:270-279."""
    if iter_step < warm_up_end:
        return iter_step / warm_up_end
    progress = (iter_step - warm_up_end) / (end_iter - warm_up_end)
    return (math.cos(math.pi * progress) + 1.0) * 0.5 * (1 - alpha) + alpha


class Trainer:
    """exp_runner.py:83-210: the per-iteration schedule around `render`, CPU."""

    def __init__(self, dataset, conf: dict, seed=0, fp16=True):
        self.ds, self.conf = dataset, conf
        torch.manual_seed(seed)
        self.np_rng = np.random.RandomState(seed)
        self.sdf = SDFNetwork(conf["encoding"], conf["sdf_network"]["d_hidden"], conf["sdf_network"]["bias"], fp16=fp16)
        self.dev = SingleVariance(conf["variance_init"])
        self.renderer = NeuSRenderer(self.sdf, self.dev, conf["gradient_method"])
        self.opt = torch.optim.Adam(list(self.sdf.parameters()) + list(self.dev.parameters()), lr=conf["learning_rate"])
        rm = conf["ray_marching"]
        self.slop = (math.log10(rm["start_step_size"]) - math.log10(rm["end_step_size"])) / conf["end_iter"]
        self.iter_step = 0
        self._set_lr()      # Runner.train calls update_learning_rate() before the loop (exp_runner.py:125): step 0 runs at lr = 0

    def _set_lr(self):
        f = lr_factor(self.iter_step, self.conf["warm_up_end"], self.conf["end_iter"], self.conf["learning_rate_alpha"])
        for g in self.opt.param_groups:
            g["lr"] = self.conf["learning_rate"] * f

    def step(self, batch=None):
        c, rm = self.conf, self.conf["ray_marching"]
        it = self.iter_step
        self.renderer.sampling_step_size = 10 ** (math.log10(rm["start_step_size"]) - self.slop * it)
        self.renderer.occupancy_grid.every_n_step(it, self.renderer.occ_eval_fn, occ_thre=rm["occ_threshold"], n=rm["occ_update_freq"])
        if it % c["increase_bindwidth_every"] == 0:
            self.sdf.increase_bandwidth()
        if batch is None:
            batch = self.ds.gen_random_patches(c["batch_size"], c["patch_size"], c["patch_size"], np_rng=self.np_rng)
        o, d, pn, Vi, nrm, msk = batch
        ps = c["patch_size"]
        near, far = self.ds.near_far_from_sphere(o[:, ps // 2, ps // 2], d[:, ps // 2, ps // 2])
        out = self.renderer.render(o, d, pn, near, far, Vi, jitter=torch.rand(o.shape[0], device=o.device))
        if out["gradients"] is None:
            self._set_lr()
            return None, out
        loss, parts = losses(out, nrm, msk, c["normal_weight"], c["mask_weight"], c["eikonal_weight"])
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        self.iter_step += 1
        self._set_lr()
        return float(loss), out
