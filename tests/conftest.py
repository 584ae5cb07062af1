import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def make_march_case(seed=0, n_rays=257, res=(128, 128, 128), step=0.01, nan_every=17, grid_kind="shell"):
    """Rays from a ring of cameras towards the unit sphere + an occupancy grid; shared by CPU/GPU tests."""
    rng = np.random.RandomState(seed)
    az = rng.uniform(0, 2 * np.pi, n_rays)
    el = rng.uniform(-0.6, 0.9, n_rays)
    o = 3.0 * np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], -1)
    target = rng.uniform(-0.7, 0.7, (n_rays, 3))
    if nan_every:
        target[::nan_every] += 2.5  # these miss the unit sphere -> NaN near/far
    d = target - o
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o, d = o.astype(np.float32), d.astype(np.float32)
    a = (d * d).sum(-1)
    b = 2 * (o * d).sum(-1)
    c = (o * o).sum(-1) - 1.0
    with np.errstate(invalid="ignore"):
        root = np.sqrt(b * b - 4 * a * c) / (2 * a)
    near = (0.5 * -b / a - root).astype(np.float32)
    far = (0.5 * -b / a + root).astype(np.float32)
    near = near + (rng.uniform(0, 1, n_rays).astype(np.float32) * np.float32(step))
    gx, gy, gz = np.meshgrid(*[(np.arange(r) + 0.5) / r * 2 - 1 for r in res], indexing="ij")
    rad = np.sqrt(gx ** 2 + gy ** 2 + gz ** 2)
    if grid_kind == "shell":
        grid = (np.abs(rad - 0.5) < 0.08) | (rng.uniform(size=res) < 0.002)
    elif grid_kind == "full":
        grid = np.ones(res, dtype=bool)
    elif grid_kind == "empty":
        grid = np.zeros(res, dtype=bool)
    else:
        grid = rng.uniform(size=res) < 0.3
    roi = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    return dict(rays_o=o, rays_d=d, t_min=near, t_max=far, roi=roi, grid=grid, step=np.float32(step))
