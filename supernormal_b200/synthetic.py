"""Synthetic DiLiGenT-MV-shaped dataset (analytic sphere) with the tensor layouts the hot
path consumes.

Mirrors the *outputs* of the reference loader -- models/dataset_loader.py:99-150
(`normals`, `masks`, `intrinsics_all(_inv)`, `pose_all`, `V_inverse_all`), :223-277
(`gen_random_patches`) and :279-297 (`near_far_from_sphere`) -- without its file I/O
(pyexr / cv2 / npz), which SURVEY.md §8 marks out of scope.  There is no network, so the
normal maps are rendered analytically: a sphere of radius 0.5 at the origin seen by pinhole
cameras on a ring (distance 3, elevation 30 deg, look-at origin), per SURVEY.md §8(d).

Conventions (same as the reference): OpenCV camera (x right, y down, z forward), `pose` is
camera-to-world, normals are world-space outward unit normals, masks are float 0/1.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import torch


@dataclass
class SyntheticScene:
    n_views: int = 20
    H: int = 512
    W: int = 612
    radius: float = 0.5
    cam_distance: float = 3.0
    elevation_deg: float = 30.0
    sphere_fraction_of_H: float = 0.6
    exclude_views: Tuple[int, ...] = (0, 4, 8, 12, 16)  # config/diligent.conf:16


def _look_at_pose(eye: np.ndarray) -> np.ndarray:
    """camera-to-world 4x4, OpenCV axes, looking at the origin, world z up."""
    fwd = -eye / np.linalg.norm(eye)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    pose = np.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, down, fwd, eye
    return pose


class SyntheticDataset:
    """Holds exactly the tensors `Dataset.__init__` leaves on the device."""

    def __init__(self, scene: SyntheticScene = SyntheticScene(), device="cpu", with_v_inverse=True):
        s = scene
        self.scene = s
        self.device = torch.device(device)
        self.n_images, self.H, self.W = s.n_views, s.H, s.W
        self.train_images = [i for i in range(s.n_views) if i not in set(s.exclude_views)]
        ang_r = math.asin(s.radius / s.cam_distance)
        focal = (0.5 * s.sphere_fraction_of_H * s.H) / math.tan(ang_r)
        K = np.eye(4)
        K[0, 0] = K[1, 1] = focal
        K[0, 2], K[1, 2] = (s.W - 1) / 2.0, (s.H - 1) / 2.0
        el = math.radians(s.elevation_deg)
        poses = []
        for v in range(s.n_views):
            az = 2.0 * math.pi * v / s.n_views
            eye = s.cam_distance * np.array([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)])
            poses.append(_look_at_pose(eye))
        dev = self.device
        self.intrinsics_all = torch.from_numpy(np.stack([K] * s.n_views)).float().to(dev)
        self.intrinsics_all_inv = torch.inverse(self.intrinsics_all).contiguous()
        self.pose_all = torch.from_numpy(np.stack(poses)).float().to(dev)
        self.focal_length = self.intrinsics_all[0][0, 0]

        normals, masks, vinv = [], [], []
        ys, xs = torch.meshgrid(torch.arange(s.H, device=dev, dtype=torch.float32),
                                torch.arange(s.W, device=dev, dtype=torch.float32), indexing="ij")
        pix = torch.stack([xs, ys, torch.ones_like(xs)], -1)  # (H,W,3)
        for v in range(s.n_views):
            Kinv = self.intrinsics_all_inv[v, :3, :3]
            R = self.pose_all[v, :3, :3]
            o = self.pose_all[v, :3, 3]
            p = pix @ Kinv.T
            d_cam = p / p.norm(dim=-1, keepdim=True)
            d = d_cam @ R.T  # world-space unit directions, (H,W,3)
            # ray-sphere hit
            b = (d * o).sum(-1)
            c = (o * o).sum() - s.radius ** 2
            disc = b * b - c
            hit = disc > 0
            t = -b - torch.sqrt(disc.clamp_min(0))
            n = (o + d * t[..., None]) / s.radius
            normals.append(torch.where(hit[..., None], n, torch.zeros_like(n)))
            masks.append(hit.float())
            if with_v_inverse:
                right = R[:, 0].expand_as(d)
                down = R[:, 1].expand_as(d)
                V = torch.stack([d, right, down], dim=-2)  # rows: ray dir, cam right, cam down
                vinv.append(torch.inverse(V).contiguous())
        self.normals = torch.stack(normals)           # [n,H,W,3]
        self.masks = torch.stack(masks)               # [n,H,W]
        self.V_inverse_all = torch.stack(vinv).contiguous() if with_v_inverse else None  # [n,H,W,3,3]
        self.object_bbox_min = np.array([-1.0, -1.0, -1.0])
        self.object_bbox_max = np.array([1.0, 1.0, 1.0])

    # ---- models/dataset_loader.py:223-277 -------------------------------------------------
    def gen_random_patches(self, num_patch: int, patch_H: int = 3, patch_W: int = 3,
                           generator: Optional[torch.Generator] = None,
                           np_rng: Optional[np.random.RandomState] = None,
                           img_idx: Optional[torch.Tensor] = None):
        dev = self.device
        cx = torch.randint(patch_W // 2, self.W - 1 - patch_W // 2, (num_patch,), device=dev, generator=generator)
        cy = torch.randint(patch_H // 2, self.H - 1 - patch_H // 2, (num_patch,), device=dev, generator=generator)
        if img_idx is None:
            choice = (np_rng or np.random).choice(self.train_images, size=[num_patch])
            img_idx = torch.as_tensor(choice, device=dev)
        return self.patches_at(img_idx.to(dev), cx, cy, patch_H, patch_W)

    def patches_at(self, img_idx, cx, cy, patch_H=3, patch_W=3):
        """Gather everything `render` + the loss need for patches centred at (cx, cy) of views img_idx."""
        dev = self.device
        ox = torch.arange(-(patch_W // 2), patch_W // 2 + 1, device=dev)
        oy = torch.arange(-(patch_H // 2), patch_H // 2 + 1, device=dev)
        px = (cx[:, None, None] + ox[None, None, :]).expand(-1, patch_H, patch_W)
        py = (cy[:, None, None] + oy[None, :, None]).expand(-1, patch_H, patch_W)
        vi = img_idx.view(-1, 1, 1).expand_as(px)
        normal = self.normals[vi, py, px]
        V_inv = self.V_inverse_all[vi, py, px]
        mask = self.masks[vi, py, px].unsqueeze(-1)
        p = torch.stack([px, py, torch.ones_like(px)], -1).float()
        p = torch.matmul(self.intrinsics_all_inv[vi, :3, :3], p[..., None])[..., 0]
        d = p / torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)
        d = torch.matmul(self.pose_all[img_idx, None, None, :3, :3], d[..., None])[..., 0]
        o = self.pose_all[img_idx, None, None, :3, 3].expand(d.shape)
        plane_n = self.pose_all[img_idx, :3, 2]
        return o.contiguous(), d.contiguous(), plane_n.contiguous(), V_inv.contiguous(), normal.contiguous(), mask.contiguous()

    # ---- models/dataset_loader.py:279-297 -------------------------------------------------
    @staticmethod
    def near_far_from_sphere(rays_o, rays_d):
        a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
        b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
        c = torch.sum(rays_o ** 2, dim=-1, keepdim=True) - 1.0
        mid = 0.5 * (-b) / a
        root = torch.sqrt(b ** 2 - 4 * a * c) / (2 * a)  # NaN for rays missing the unit sphere
        return (mid - root)[..., 0], (mid + root)[..., 0]


# diligent.conf / own_objects.conf as plain dicts (pyhocon is not installed)
DILIGENT_CONF = dict(
    learning_rate=5e-4, learning_rate_alpha=0.05, end_iter=5000, increase_bindwidth_every=350,
    gradient_method="dfd", batch_size=2048, patch_size=3, warm_up_end=50, loss_type="l2",
    normal_weight=1.0, eikonal_weight=1.0, mask_weight=1.0,
    sdf_network=dict(d_out=1, d_in=3, d_hidden=64, n_layers=1, bias=0.6, geometric_init=True,
                     weight_norm=True, input_concat=True),
    variance_init=0.5,
    ray_marching=dict(start_step_size=1e-2, end_step_size=1e-3, occ_threshold=0.1, occ_update_freq=8),
    encoding=dict(otype="HashGrid", n_levels=14, n_features_per_level=2, log2_hashmap_size=19,
                  base_resolution=32, per_level_scale=1.3195079107728942),
)
OWN_OBJECTS_CONF = dict(DILIGENT_CONF, end_iter=30000, increase_bindwidth_every=2000, warm_up_end=500,
                        sdf_network=dict(DILIGENT_CONF["sdf_network"], bias=0.8))
