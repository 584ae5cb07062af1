"""nerfacc-0.3.5-shaped operator API (boundary A of SURVEY.md §8b) on top of libsnb200.

Exactly the names models/renderer.py:5-7 imports, with the same argument meaning, shapes, dtypes
and error behaviour as the reference wrappers (NA = third_parties/nerfacc-0.3.5/nerfacc-0.3.5/nerfacc):
ContractionType (NA/contraction.py:12-62), OccupancyGrid (NA/grid.py:113-294), ray_marching
(NA/ray_marching.py:14-222), pack_info (NA/pack.py:47-77), render_visibility /
render_transmittance_from_alpha / render_weight_from_alpha[_patch_based] /
accumulate_along_rays[_patch_based] (NA/vol_rendering.py).  CUDA tensors only, like the reference.
"""
from __future__ import annotations

from enum import Enum
from typing import Callable, List, Optional, Tuple, Union

import torch
from torch import Tensor

from . import _lib
from ._lib import call, ptr


class ContractionType(Enum):
    """NA/contraction.py:12-62. Only AABB is on SuperNormal's path (models/renderer.py:46)."""
    AABB = 0
    UN_BOUNDED_TANH = 1
    UN_BOUNDED_SPHERE = 2

    def to_cpp_version(self):
        return self.value


def _require_cuda(t: Tensor):
    if not t.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")


# ------------------------------------------------------------------------------------------
# packing
# ------------------------------------------------------------------------------------------
@torch.no_grad()
def packed_info_from_counts(num_steps: Tensor) -> Tuple[Tensor, Tensor]:
    """(offset,count) pairs + device total from per-ray counts, no host sync."""
    n = num_steps.numel()
    packed = torch.empty((n, 2), dtype=torch.int32, device=num_steps.device)
    total = torch.zeros(1, dtype=torch.int32, device=num_steps.device)
    call("snb_packed_info_from_counts", n, ptr(num_steps), ptr(packed), ptr(total))
    return packed, total


@torch.no_grad()
def pack_info(ray_indices: Tensor, n_rays: Optional[int] = None) -> Tensor:
    """NA/pack.py:47-77."""
    assert ray_indices.dim() == 1, "ray_indices must be a 1D tensor with shape (n_samples)."
    _require_cuda(ray_indices)
    ray_indices = ray_indices.contiguous().long()
    if n_rays is None:
        mx = torch.empty(1, dtype=torch.int64, device=ray_indices.device)
        call("snb_max_i64", ray_indices.numel(), ptr(ray_indices), ptr(mx))
        n_rays = int(mx.item()) + 1  # the reference syncs here too (NA/pack.py:67)
    num = torch.empty(n_rays, dtype=torch.int32, device=ray_indices.device)
    call("snb_count_by_ray", ray_indices.numel(), ptr(ray_indices), n_rays, ptr(num))
    return packed_info_from_counts(num)[0]


# ------------------------------------------------------------------------------------------
# rendering weights / transmittance / accumulation
# ------------------------------------------------------------------------------------------
class _WeightFromAlphaPatch(torch.autograd.Function):
    """NA/vol_rendering.py:990-1010 (_RenderingWeightFromAlphaPatchBasedNaive); P=1 is
    _RenderingWeightFromAlphaNaive (:968-988)."""

    @staticmethod
    def forward(ctx, packed_info, alphas):
        packed_info = packed_info.contiguous()
        alphas = alphas.contiguous()
        P = alphas.shape[1] if alphas.dim() == 3 else 1
        weights = torch.empty_like(alphas)
        call("snb_weight_from_alpha_patch_fwd", packed_info.shape[0], P, ptr(packed_info), ptr(alphas), ptr(weights))
        if ctx.needs_input_grad[1]:
            ctx.save_for_backward(packed_info, alphas, weights)
        ctx.P = P
        return weights

    @staticmethod
    def backward(ctx, grad_weights):
        grad_weights = grad_weights.contiguous()
        packed_info, alphas, weights = ctx.saved_tensors
        grad_alphas = torch.empty_like(alphas)
        call("snb_weight_from_alpha_patch_bwd", packed_info.shape[0], ctx.P, ptr(packed_info), ptr(alphas),
             ptr(weights), ptr(grad_weights), ptr(grad_alphas))
        return None, grad_alphas


def render_weight_from_alpha_patch_based(alphas: Tensor, ray_indices: Tensor, *, n_rays: Optional[int] = None) -> Tensor:
    """NA/vol_rendering.py:533-576. alphas (n_samples, patch_size, 1)."""
    _require_cuda(alphas)
    packed_info = pack_info(ray_indices, n_rays=n_rays)
    return _WeightFromAlphaPatch.apply(packed_info, alphas)


@torch.no_grad()
def render_transmittance_from_alpha(alphas: Tensor, *, ray_indices: Optional[Tensor] = None,
                                    packed_info: Optional[Tensor] = None, n_rays: Optional[int] = None) -> Tensor:
    """Exclusive product of (1-alpha) along each ray, serial order (the reference's naive kernel,
    CS/render_transmittance.cu:85-112; its CUB route differs only in multiplication order)."""
    assert ray_indices is not None or packed_info is not None, "Either ray_indices or packed_info should be provided."
    _require_cuda(alphas)
    if packed_info is None:
        packed_info = pack_info(ray_indices, n_rays=n_rays)
    alphas = alphas.contiguous()
    T = torch.empty_like(alphas)
    call("snb_transmittance_from_alpha", packed_info.shape[0], ptr(packed_info.contiguous()), ptr(alphas), ptr(T))
    return T


@torch.no_grad()
def render_visibility(alphas: Tensor, *, ray_indices: Optional[Tensor] = None, packed_info: Optional[Tensor] = None,
                      n_rays: Optional[int] = None, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0) -> Tensor:
    """NA/vol_rendering.py:680-748."""
    T = render_transmittance_from_alpha(alphas, ray_indices=ray_indices, packed_info=packed_info, n_rays=n_rays)
    vis = T >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis.squeeze(-1)


def render_weight_from_alpha(alphas: Tensor, *, ray_indices: Optional[Tensor] = None,
                             packed_info: Optional[Tensor] = None, n_rays: Optional[int] = None) -> Tensor:
    """NA/vol_rendering.py:624-677. alphas (n_samples, 1)."""
    assert ray_indices is not None or packed_info is not None, "Either ray_indices or packed_info should be provided."
    _require_cuda(alphas)
    if packed_info is None:
        packed_info = pack_info(ray_indices, n_rays=n_rays)
    return _WeightFromAlphaPatch.apply(packed_info, alphas)


class _Accumulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, values, ray_indices, n_out):
        P = weights.shape[1] if weights.dim() == 3 else 1
        D = values.shape[-1] if values is not None else weights.shape[-1]
        w = weights.contiguous()
        v = values.contiguous() if values is not None else None
        idx = ray_indices.contiguous().long()
        shape = (n_out, P, D) if weights.dim() == 3 else (n_out, D)
        out = torch.empty(shape, dtype=w.dtype, device=w.device)
        call("snb_accumulate_fwd", w.shape[0], P, D, ptr(w), ptr(v), ptr(idx), n_out, ptr(out))
        ctx.save_for_backward(w, v, idx)
        ctx.dims = (P, D)
        return out

    @staticmethod
    def backward(ctx, go):
        w, v, idx = ctx.saved_tensors
        P, D = ctx.dims
        go = go.contiguous()
        gw = torch.empty_like(w) if ctx.needs_input_grad[0] else None
        gv = torch.empty_like(v) if (v is not None and ctx.needs_input_grad[1]) else None
        if gw is not None or gv is not None:
            call("snb_accumulate_bwd", w.shape[0], P, D, ptr(w), ptr(v), ptr(idx), ptr(go), ptr(gw), ptr(gv))
        return gw, gv, None, None


def accumulate_along_rays_patch_based(weights: Tensor, ray_indices: Tensor, values: Optional[Tensor] = None,
                                      n_patches: Optional[int] = None) -> Tensor:
    """NA/vol_rendering.py:269-335."""
    assert ray_indices.dim() == 1 and weights.dim() == 3
    _require_cuda(weights)
    if values is not None:
        assert values.dim() == 3 and values.shape[0] == weights.shape[0], \
            "Invalid shapes: {} vs {}".format(values.shape, weights.shape)
    D = values.shape[-1] if values is not None else weights.shape[-1]
    if ray_indices.numel() == 0:
        assert n_patches is not None
        return torch.zeros((n_patches, weights.shape[1], D), device=weights.device)
    if n_patches is None:
        n_patches = int(ray_indices.max()) + 1
    return _Accumulate.apply(weights, values, ray_indices, n_patches)


def accumulate_along_rays(weights: Tensor, ray_indices: Tensor, values: Optional[Tensor] = None,
                          n_rays: Optional[int] = None) -> Tensor:
    """NA/vol_rendering.py:132-198."""
    assert ray_indices.dim() == 1 and weights.dim() == 2
    _require_cuda(weights)
    if values is not None:
        assert values.dim() == 2 and values.shape[0] == weights.shape[0], \
            "Invalid shapes: {} vs {}".format(values.shape, weights.shape)
    D = values.shape[-1] if values is not None else weights.shape[-1]
    if ray_indices.numel() == 0:
        assert n_rays is not None
        return torch.zeros((n_rays, D), device=weights.device)
    if n_rays is None:
        n_rays = int(ray_indices.max()) + 1
    return _Accumulate.apply(weights, values, ray_indices, n_rays)


# ------------------------------------------------------------------------------------------
# occupancy grid
# ------------------------------------------------------------------------------------------
class OccupancyGrid(torch.nn.Module):
    """NA/grid.py:113-294.  Buffers `_roi_aabb`, `_binary`, `resolution`, `occs` keep the reference's
    names (state-dict compatible); the 50 MB `grid_coords` / 17 MB `grid_indices` helper buffers are
    not materialised -- cell coordinates are derived from the cell index inside the kernel."""

    NUM_DIM: int = 3

    def __init__(self, roi_aabb: Union[List[int], Tensor], resolution: Union[int, List[int], Tensor] = 128,
                 contraction_type: ContractionType = ContractionType.AABB) -> None:
        super().__init__()
        if isinstance(resolution, int):
            resolution = [resolution] * self.NUM_DIM
        if isinstance(resolution, (list, tuple)):
            resolution = torch.tensor(resolution, dtype=torch.int32)
        assert isinstance(resolution, Tensor), f"Invalid type: {type(resolution)}"
        assert resolution.shape == (self.NUM_DIM,), f"Invalid shape: {resolution.shape}"
        if isinstance(roi_aabb, (list, tuple)):
            roi_aabb = torch.tensor(roi_aabb, dtype=torch.float32)
        assert isinstance(roi_aabb, Tensor), f"Invalid type: {type(roi_aabb)}"
        assert roi_aabb.shape == torch.Size([self.NUM_DIM * 2]), f"Invalid shape: {roi_aabb.shape}"
        if contraction_type != ContractionType.AABB:
            raise NotImplementedError("supernormal_b200 implements the AABB contraction only (models/renderer.py:46)")
        self._res = [int(r) for r in resolution.tolist()]
        self.num_cells = int(resolution.prod().item())
        self.register_buffer("_roi_aabb", roi_aabb.clone())
        self.register_buffer("_binary", torch.zeros(self._res, dtype=torch.bool))
        self._contraction_type = contraction_type
        self.register_buffer("resolution", resolution)
        self.register_buffer("occs", torch.zeros(self.num_cells))
        self.register_buffer("_ws", torch.zeros(1, dtype=torch.float64), persistent=False)

    @property
    def roi_aabb(self) -> Tensor:
        return self._roi_aabb

    @property
    def binary(self) -> Tensor:
        return self._binary

    @property
    def contraction_type(self) -> ContractionType:
        return self._contraction_type

    @property
    def device(self) -> torch.device:
        return self._roi_aabb.device

    @torch.no_grad()
    def _sample_uniform_and_occupied_cells(self, n: int) -> Tensor:
        """NA/grid.py:182-194."""
        uniform_indices = torch.randint(self.num_cells, (n,), device=self.device)
        occupied_indices = torch.nonzero(self._binary.flatten())[:, 0]
        if n < len(occupied_indices):
            selector = torch.randint(len(occupied_indices), (n,), device=self.device)
            occupied_indices = occupied_indices[selector]
        return torch.cat([uniform_indices, occupied_indices], dim=0)

    @torch.no_grad()
    def _update(self, step: int, occ_eval_fn: Callable, occ_thre: float = 0.01, ema_decay: float = 0.95,
                warmup_steps: int = 256, rand: Optional[Tensor] = None, indices: Optional[Tensor] = None) -> None:
        """NA/grid.py:197-239.  `rand` / `indices` are injectable for parity tests."""
        _require_cuda(self.occs)
        if indices is None and step >= warmup_steps:
            indices = self._sample_uniform_and_occupied_cells(self.num_cells // 4)
        n = self.num_cells if indices is None else indices.numel()
        if indices is not None:
            indices = indices.contiguous().long()
        if rand is None:
            rand = torch.rand((n, 3), device=self.device)
        x = torch.empty((n, 3), device=self.device)
        call("snb_occgrid_points", n, ptr(indices), ptr(rand.contiguous()), *self._res, ptr(self._roi_aabb), ptr(x))
        occ = occ_eval_fn(x).squeeze(-1).contiguous().float()
        call("snb_occgrid_ema", n, ptr(indices), ptr(occ), float(ema_decay), ptr(self.occs))
        binary = torch.empty(self._res, dtype=torch.bool, device=self.device)
        call("snb_occgrid_binarize", self.num_cells, ptr(self.occs), float(occ_thre), ptr(binary), ptr(self._ws))
        self._binary = binary

    @torch.no_grad()
    def every_n_step(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2, ema_decay: float = 0.95,
                     warmup_steps: int = 256, n: int = 16) -> None:
        """NA/grid.py:242-277."""
        if not self.training:
            raise RuntimeError(
                "You should only call this function only during training. "
                "Please call _update() directly if you want to update the "
                "field during inference.")
        if step % n == 0 and self.training:
            self._update(step=step, occ_eval_fn=occ_eval_fn, occ_thre=occ_thre, ema_decay=ema_decay,
                         warmup_steps=warmup_steps)


# ------------------------------------------------------------------------------------------
# ray marching
# ------------------------------------------------------------------------------------------
@torch.no_grad()
def _march(rays_o, rays_d, t_min, t_max, roi, binary, step_size, cone_angle):
    """The native part of NA/ray_marching.py:177-190 (`_C.ray_marching`, CS/ray_marching.cu:194-289).
    Returns (packed_info i32[n,2], ray_indices i64[S], t_starts f32[S,1], t_ends f32[S,1])."""
    n = rays_o.shape[0]
    dev = rays_o.device
    assert rays_o.dim() == 2 and rays_o.shape[1] == 3 and rays_d.shape == rays_o.shape
    assert t_min.dim() == 1 and t_max.dim() == 1 and roi.numel() == 6 and binary.dim() == 3
    res = list(binary.shape)
    grid_u8 = binary.view(torch.uint8) if binary.dtype == torch.bool else binary.to(torch.uint8)
    args = (n, ptr(rays_o), ptr(rays_d), ptr(t_min), ptr(t_max), ptr(roi), *res, ptr(grid_u8), float(step_size), float(cone_angle))
    counts = torch.empty(n, dtype=torch.int32, device=dev)
    call("snb_march_count", *args, ptr(counts))
    packed, total = packed_info_from_counts(counts)
    S = int(total.item())  # same single host sync as the reference (CS/ray_marching.cu:261)
    t0 = torch.empty((S, 1), device=dev)
    t1 = torch.empty((S, 1), device=dev)
    ridx = torch.empty((S,), dtype=torch.int64, device=dev)
    call("snb_march_emit", *args, ptr(packed), S, ptr(ridx), None, ptr(t0), ptr(t1))
    return packed, ridx, t0, t1


@torch.no_grad()
def ray_marching(rays_o: Tensor, rays_d: Tensor, t_min: Optional[Tensor] = None, t_max: Optional[Tensor] = None,
                 scene_aabb: Optional[Tensor] = None, grid=None, sigma_fn: Optional[Callable] = None,
                 alpha_fn: Optional[Callable] = None, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0,
                 near_plane: Optional[float] = None, far_plane: Optional[float] = None,
                 render_step_size: float = 1e-3, stratified: bool = False, cone_angle: float = 0.0):
    """NA/ray_marching.py:14-222 -> (ray_indices i64[S], t_starts f32[S,1], t_ends f32[S,1])."""
    if not rays_o.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    if alpha_fn is not None and sigma_fn is not None:
        raise ValueError("Only one of `alpha_fn` and `sigma_fn` should be provided.")
    if t_min is None or t_max is None:
        if scene_aabb is not None:
            # NA/intersection.py is off SuperNormal's path: it always passes t_min/t_max.
            raise NotImplementedError("ray_aabb_intersect is out of scope: pass t_min and t_max")
        t_min = torch.zeros_like(rays_o[..., 0])
        t_max = torch.ones_like(rays_o[..., 0]) * 1e10
    if near_plane is not None:
        t_min = torch.clamp(t_min, min=near_plane)
    if far_plane is not None:
        t_max = torch.clamp(t_max, max=far_plane)
    if stratified:
        t_min = t_min + torch.rand_like(t_min) * render_step_size
    if grid is not None:
        roi, binary = grid.roi_aabb, grid.binary
        if grid.contraction_type != ContractionType.AABB:
            raise NotImplementedError("AABB contraction only")
    else:
        roi = torch.tensor([-1e10, -1e10, -1e10, 1e10, 1e10, 1e10], dtype=torch.float32, device=rays_o.device)
        binary = torch.ones([1, 1, 1], dtype=torch.bool, device=rays_o.device)
    packed_info, ray_indices, t_starts, t_ends = _march(
        rays_o.contiguous(), rays_d.contiguous(), t_min.contiguous(), t_max.contiguous(),
        roi.contiguous(), binary.contiguous(), render_step_size, cone_angle)
    if sigma_fn is not None or alpha_fn is not None:
        if sigma_fn is not None:
            sigmas = sigma_fn(t_starts, t_ends, ray_indices)
            assert sigmas.shape == t_starts.shape, "sigmas must have shape of (N, 1)! Got {}".format(sigmas.shape)
            alphas = 1.0 - torch.exp(-sigmas * (t_ends - t_starts))
        else:
            alphas = alpha_fn(t_starts, t_ends, ray_indices)
            assert alphas.shape == t_starts.shape, "alphas must have shape of (N, 1)! Got {}".format(alphas.shape)
        masks = render_visibility(alphas, ray_indices=ray_indices, packed_info=packed_info,
                                  early_stop_eps=early_stop_eps, alpha_thre=alpha_thre, n_rays=rays_o.shape[0])
        ray_indices, t_starts, t_ends = ray_indices[masks], t_starts[masks], t_ends[masks]
    return ray_indices, t_starts, t_ends


__all__ = ["ContractionType", "OccupancyGrid", "ray_marching", "pack_info", "render_visibility",
           "render_transmittance_from_alpha", "render_weight_from_alpha", "render_weight_from_alpha_patch_based",
           "accumulate_along_rays", "accumulate_along_rays_patch_based"]
