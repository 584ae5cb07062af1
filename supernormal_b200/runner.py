"""Thin harness around the hot path: the eval renderer of
exp_runner.py:397-481 (validate_normal_patch_based) and :526-577 (eval_mae), and the time-to-mesh driver
(Runner.train + extract_geometry, exp_runner.py:147-246,483-506).  Host logic only; all tensor math is in libsnb200.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Dict, Optional

import numpy as np
import torch

from . import mesh, mesh_post
from ._lib import call, ptr
from .trainer import FusedTrainer, P, make_batch_struct


@torch.no_grad()
def render_patches(tr: FusedTrainer, batch: dict, step_size: float, jitter: Optional[torch.Tensor] = None):
    """renderer.render(..., mode='eval') (models/renderer.py:63-276 without backward): comp_normal [N,3,3,3],
    weight_sum [N,3,3,1] for N <= tr.n_patches patches."""
    m, b = tr.model, tr.buf
    n = batch["rays_d"].shape[0]
    assert n <= tr.n_patches, "eval batch exceeds the trainer's buffer capacity"
    zeros = lambda *s: torch.zeros(*s, device=tr.device)
    bs = make_batch_struct(batch["rays_o"], batch["rays_d"], batch["plane_n"], batch["near"], batch["far"], batch["v_inv"],
                           batch.get("normal_gt", zeros(n, P, 3)), batch.get("mask", zeros(n, P)))
    keep = [bs]
    m.prep(batch.get("mask", zeros(n, P)), b.stats)
    net = m.net_struct()
    rb, rn, rs = C.byref(bs), C.byref(net), C.byref(b.struct)
    g = tr.grid
    call("snb_march_visible", rb, rn, ptr(g.roi_aabb), *g._res, ptr(g.binary.view(torch.uint8)), float(step_size), ptr(jitter), 1e-8, rs)
    call("snb_compact_samples", n, rs)
    call("snb_sdf_fwd_patch", rb, rn, rs, ptr(b.sdf), ptr(b.feats))
    call("snb_render_fused", rb, rn, rs, ptr(b.sdf), 0.0, 0.0, 0.0, ptr(b.comp), ptr(b.wsum), None, None, ptr(b.stats))
    del keep
    return b.comp[:n].view(n, 3, 3, 3).clone(), b.wsum[:n].view(n, 3, 3, 1).clone()


@torch.no_grad()
def render_normal_pixel_based(tr: FusedTrainer, rays_o: torch.Tensor, rays_d: torch.Tensor, near: torch.Tensor, far: torch.Tensor,
                              step_size: Optional[float] = None, stratified: bool = True) -> Dict[str, torch.Tensor]:
    """NeuSRenderer.render_normal_pixel_based (models/renderer.py:278-351): per-pixel validation render.  Per-ray marching through the
    occupancy grid with the NeuS alpha of the centre-ray form as alpha_fn (alpha_thre = 0, early_stop_eps = 1e-3), alpha again on the
    surviving samples, analytic normals SDFNetwork.gradient at the interval midpoints (snb_sdf_eval_grad: SDF + gradient in one
    pass, no autograd graph), render_weight_from_alpha and accumulate_along_rays (the P = 1 kernels of rows a10 / a11).
    -> comp_normal [n,3], comp_depth [n,1] (the reference's return values) + weight_sum [n,1], n_samples."""
    from . import nerfacc_api as na
    m = tr.model
    m.prep()
    inv_s = m.net[2369]                       # clip(exp(10 variance), 1e-6, 1e6), folded by snb_prep_net (sdf_core.cuh:kOffInvS)
    rays_o, rays_d = rays_o.contiguous().float(), rays_d.contiguous().float()

    def alpha_fn(t_starts, t_ends, ray_indices):
        ridx = ray_indices.long()
        o, d = rays_o[ridx], rays_d[ridx]
        ps, pe = o + d * t_starts, o + d * t_ends
        nxt = torch.cat([t_starts[1:], t_starts[-1:]], 0)
        diff = ((t_ends - nxt) != 0).squeeze(-1)                    # models/renderer.py:288-293
        sdf_all = m.sdf(torch.cat([ps, pe[diff].reshape(-1, 3)], 0))
        s0 = sdf_all[: ps.shape[0]]
        s1 = torch.cat([s0[1:], s0[-1:]], 0)
        s1[diff] = sdf_all[ps.shape[0]:]
        c, n = torch.sigmoid(s0 * inv_s), torch.sigmoid(s1 * inv_s)
        return ((c - n + 1e-5) / (c + 1e-5)).view(-1).clip(0.0, 1.0).reshape(-1, 1)

    step = tr.step_size(tr.iter_step) if step_size is None else step_size
    ridx, t0, t1 = na.ray_marching(rays_o, rays_d, t_min=near.reshape(-1), t_max=far.reshape(-1), grid=tr.grid, render_step_size=step,
                                   stratified=stratified, cone_angle=0.0, alpha_thre=0.0, early_stop_eps=1e-3, alpha_fn=alpha_fn)
    n = rays_o.shape[0]
    if ridx.numel() == 0:
        z = torch.zeros(n, 1, device=rays_o.device)
        return {"comp_normal": torch.zeros(n, 3, device=rays_o.device), "comp_depth": z, "weight_sum": z.clone(), "n_samples": 0}
    alpha = alpha_fn(t0, t1, ridx)
    mid = (t0 + t1) / 2.0
    _, grad = m.sdf_and_gradient(rays_o[ridx.long()] + rays_d[ridx.long()] * mid)
    w = na.render_weight_from_alpha(alpha, ray_indices=ridx, n_rays=n)
    return {"comp_normal": na.accumulate_along_rays(w, ridx, values=grad, n_rays=n),
            "comp_depth": na.accumulate_along_rays(w, ridx, values=mid, n_rays=n),
            "weight_sum": na.accumulate_along_rays(w, ridx, values=None, n_rays=n), "n_samples": int(ridx.numel())}


@torch.no_grad()
def validate_normal_patch_based(tr: FusedTrainer, idx: int, eval_patch_size: int = 1024, stratified: bool = True) -> torch.Tensor:
    """World-space rendered normal map [3*(H//3), 3*(W//3), 3] of view idx from non-overlapping 3x3 patches
    (Dataset.gen_patches_at, models/dataset_loader.py:177-221; tiling loop of exp_runner.py:423-453)."""
    ds = tr.ds
    dev = tr.device
    ny, nx = ds.H // 3, ds.W // 3
    cy, cx = torch.meshgrid(torch.arange(ny, device=dev) * 3 + 1, torch.arange(nx, device=dev) * 3 + 1, indexing="ij")
    cy, cx = cy.reshape(-1), cx.reshape(-1)
    out = torch.zeros(ny * nx, 3, 3, 3, device=dev)
    step = tr.step_size(tr.iter_step)
    for s in range(0, ny * nx, eval_patch_size):
        e = min(s + eval_patch_size, ny * nx)
        img = torch.full((e - s,), idx, device=dev, dtype=torch.long)
        o, d, pn, vinv, nrm, msk = ds.patches_at(img, cx[s:e], cy[s:e], 3, 3)
        near, far = ds.near_far_from_sphere(o[:, 1, 1], d[:, 1, 1])
        batch = dict(rays_o=o[:, 1, 1].contiguous(), rays_d=d.view(-1, P, 3), plane_n=pn, near=near.contiguous(), far=far.contiguous(),
                     v_inv=vinv.view(-1, P, 9))
        jit = torch.rand(e - s, device=dev) if stratified else None
        out[s:e] = render_patches(tr, batch, step, jit)[0]
    return out.view(ny, nx, 3, 3, 3).permute(0, 2, 1, 3, 4).reshape(ny * 3, nx * 3, 3)


@torch.no_grad()
def eval_mae(tr: FusedTrainer) -> Dict[str, float]:
    """Mean angular error in degrees over masked pixels: all views / held-out views (exp_runner.py:526-577)."""
    ds = tr.ds
    errs, errs_test = [], []
    for idx in range(ds.n_images):
        nm = validate_normal_patch_based(tr, idx)
        nm = nm / (1e-10 + nm.norm(dim=-1, keepdim=True))
        h, w = nm.shape[:2]
        gt, mask = ds.normals[idx, :h, :w], ds.masks[idx, :h, :w] > 0.5
        ae = torch.rad2deg(torch.arccos((gt * nm).sum(-1).clamp(-1, 1)))[mask]
        errs.append(ae)
        if idx not in ds.train_images:
            errs_test.append(ae)
    out = {"mae_allview": float(torch.cat(errs).mean())}
    if errs_test:
        out["mae_testview"] = float(torch.cat(errs_test).mean())
    return out


@torch.no_grad()
def validate_mesh(tr: FusedTrainer, resolution: int = 512, threshold: float = 0.0):
    """Runner.validate_mesh (exp_runner.py:483-506) without the file I/O: extract_geometry on the object bounding box, then
    remove_isolated_clusters (largest edge-connected cluster).  -> (vertices float64 [V,3], triangles int32 [T,3]); None on
    ranks > 0 of a data-parallel run."""
    res = mesh.extract_geometry(tr.model, tr.ds.object_bbox_min, tr.ds.object_bbox_max, resolution, threshold)
    if res is None:
        return None
    return mesh_post.remove_isolated_clusters(*res)


@torch.no_grad()
def find_visible_points(tr: FusedTrainer, resolution_level: int = 1) -> np.ndarray:
    """Runner.find_visible_points (exp_runner.py:580-592): the surface point behind every foreground pixel of every view, by sphere
    tracing the SDF (mesh_post.find_visible_points) -- the points Chamfer distance / F-score are computed on (:573-577)."""
    tr.model.prep()
    return mesh_post.find_visible_points(tr.ds, lambda x: tr.model.sdf(x), resolution_level=resolution_level).cpu().numpy()


def chamfer_distance_and_f1_score(ref_points: np.ndarray, eval_points: np.ndarray, f_threshold: float = 0.5):
    """models/cd_and_fscore.py:5-29 (symmetric mean nearest-neighbour distance / 2; F-score at f_threshold)."""
    from scipy.spatial import KDTree
    d_e2r, _ = KDTree(ref_points).query(eval_points, k=1, p=2)
    d_r2e, _ = KDTree(eval_points).query(ref_points, k=1, p=2)
    cd = (np.mean(d_e2r) + np.mean(d_r2e)) / 2
    precision, recall = np.mean(d_e2r < f_threshold), np.mean(d_r2e < f_threshold)
    return float(cd), float(2 * precision * recall / (precision + recall))


def evaluate_sphere_mesh(v: np.ndarray, t: np.ndarray, radius: float, world_scale: float = 100.0) -> dict:
    """Chamfer distance / F-score (models/cd_and_fscore.py:5-29) of a mesh against the analytic sphere, both sampled
    with 200 000 points (area-weighted on the mesh); world_scale maps normalised units to 'mm'."""
    rng = np.random.RandomState(0)
    gt = rng.randn(200000, 3)
    gt = gt / np.linalg.norm(gt, axis=1, keepdims=True) * radius * world_scale
    p = v[t.astype(np.int64)] * world_scale
    area = 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1)
    pick = rng.choice(len(t), 200000, p=area / area.sum())
    uvw = rng.dirichlet([1, 1, 1], 200000)
    ev = (p[pick] * uvw[:, :, None]).sum(1)
    cd, f = chamfer_distance_and_f1_score(gt, ev, 0.5)
    return dict(chamfer_mm=cd, fscore=f, radius_mean=float(np.linalg.norm(v, axis=1).mean()), radius_std=float(np.linalg.norm(v, axis=1).std()))


def time_to_mesh(dataset, conf: dict, resolution: int = 512, device="cuda", seed: int = 0, world_scale: float = 100.0,
                 evaluate: bool = True) -> dict:
    """Train conf['end_iter'] iterations from random init, then extract the mesh (BASELINE.json 'time-to-mesh').
    The synthetic scene is a sphere of radius scene.radius; world_scale maps normalised units to 'mm' so that the
    reference's 0.5 mm F-score threshold (models/cd_and_fscore.py:5) keeps its meaning (radius 0.5 -> 50 mm)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    tr = FusedTrainer(dataset, conf, device=device, seed=seed, world_size=world, rank=rank)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(int(conf["end_iter"])):
        tr.train_step()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    res = mesh.extract_geometry(tr.model, dataset.object_bbox_min, dataset.object_bbox_max, resolution, 0.0)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out = {"train_s": t1 - t0, "mesh_s": t2 - t1, "time_to_mesh_s": t2 - t0, "iters": int(conf["end_iter"]), "resolution": resolution,
           "world_size": world, "loss": tr.loss_terms()}
    if res is None:
        return out
    v, t = res
    out.update(n_vertices=int(v.shape[0]), n_triangles=int(t.shape[0]))
    if evaluate and rank == 0:
        out.update(evaluate_sphere_mesh(v, t, float(dataset.scene.radius), world_scale))
        out.update(eval_mae(tr))
    out["vertices"], out["triangles"] = v, t
    return out
