"""The fused marcher (snb_march_visible: speculative empty-space windows, closed-form add chains, in-warp visibility)
must emit, with the visibility cut disabled, EXACTLY the samples of the drop-in marcher (snb_march_count/emit), which is
itself bit-exact against the reference kernel (tests/test_ref_differential.py) -- across step sizes from the start
(1e-2) to the end (1e-3) of the schedule, binade-crossing t ranges, NaN near/far and degenerate grids."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import make_march_case

pytestmark = pytest.mark.gpu

ENC = dict(otype="HashGrid", n_levels=4, n_features_per_level=2, log2_hashmap_size=12, base_resolution=8, per_level_scale=1.5)


def _run_fused(cuda, c, step, cap):
    from supernormal_b200 import nerfacc_api as na
    from supernormal_b200._lib import call, ptr
    from supernormal_b200.trainer import SDFModel, SampleBuffers, make_batch_struct
    n = c["rays_o"].shape[0]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    rays_o, d_c = t(c["rays_o"]), t(c["rays_d"])
    rays_d = d_c[:, None, :].repeat(1, 9, 1).contiguous()
    near, far, roi, grid = t(c["t_min"]), t(c["t_max"]), t(c["roi"]), t(c["grid"])
    model = SDFModel(ENC, device=cuda)
    model.n_active = 2
    model.prep()
    buf = SampleBuffers(n, cap, cap, 4, cuda)
    bs = make_batch_struct(rays_o, rays_d, torch.zeros(n, 3, device=cuda), near, far, None, None, None)
    net = model.net_struct()
    call("snb_march_visible", C.byref(bs), C.byref(net), ptr(roi), *grid.shape, ptr(grid.view(torch.uint8)), float(step), None, 0.0,
         C.byref(buf.struct))
    call("snb_compact_samples", n, C.byref(buf.struct))
    S, _, overflow, _ = buf.totals.tolist()
    assert overflow == 0
    packed_ref, ridx_ref, t0_ref, t1_ref = na._march(rays_o, d_c.contiguous(), near, far, roi, grid, float(step), 0.0)
    return buf, S, packed_ref, ridx_ref, t0_ref, t1_ref


@pytest.mark.parametrize("kind,step,seed", [("shell", 1e-2, 0), ("shell", 3.1e-3, 1), ("shell", 1.0471e-3, 2), ("random", 1e-3, 3),
                                             ("random", 7.3e-3, 4), ("full", 2e-3, 5), ("empty", 1e-3, 6)])
def test_fused_marcher_samples_equal_dropin_marcher(cuda, kind, step, seed):
    step = float(np.float32(step))
    c = make_march_case(seed=seed, n_rays=700, step=step, grid_kind=kind)
    cap = int(2.0 / step) + 64
    buf, S, packed_ref, ridx_ref, t0_ref, t1_ref = _run_fused(cuda, c, step, cap)
    assert S == ridx_ref.numel()
    assert torch.equal(buf.packed_info, packed_ref)
    assert torch.equal(buf.patch_idx[:S].long(), ridx_ref)
    assert torch.equal(buf.t0[:S].view(torch.int32), t0_ref[:, 0].view(torch.int32))
    assert torch.equal(buf.t1[:S].view(torch.int32), t1_ref[:, 0].view(torch.int32))
    if kind != "empty":
        assert S > 0


@pytest.mark.parametrize("scale,step", [(0.35, 2e-3), (0.66, 1.3e-3), (1.4, 4e-3)])
def test_fused_marcher_across_binades(cuda, scale, step):
    """Cameras at other distances put t in other binades ([0.5,1), [1,2), [4,8)) and make runs cross binade boundaries
    (t passing 1.0, 2.0, 4.0): the closed-form add chains must hand over to the serial replay exactly there."""
    step = float(np.float32(step))
    c = make_march_case(seed=11, n_rays=600, step=step, grid_kind="random", nan_every=0)
    o = c["rays_o"] * np.float32(scale)          # ring of cameras at distance 3*scale (inside the unit sphere when < 1)
    d = c["rays_d"]
    b = (o * d).sum(-1)
    disc = b * b - ((o * o).sum(-1) - np.float32(3.0))
    far = (-b + np.sqrt(np.maximum(disc, 0))).astype(np.float32)          # exit of the radius-sqrt(3) ball: covers the whole grid
    near = np.maximum(-b - np.sqrt(np.maximum(disc, 0)), np.float32(0.05)).astype(np.float32)
    c.update(rays_o=o.astype(np.float32), t_min=near, t_max=far)
    cap = int(4.0 / step) + 64
    buf, S, packed_ref, ridx_ref, t0_ref, t1_ref = _run_fused(cuda, c, step, cap)
    assert S == ridx_ref.numel() and S > 0
    assert torch.equal(buf.packed_info, packed_ref)
    assert torch.equal(buf.t0[:S].view(torch.int32), t0_ref[:, 0].view(torch.int32))
    assert torch.equal(buf.t1[:S].view(torch.int32), t1_ref[:, 0].view(torch.int32))
    tt = t0_ref[:, 0]
    assert tt.min() < 1.0 or tt.max() > 4.0 or scale > 1   # the case really leaves the [2,4) binade
