"""Loads the reference's OWN `models/renderer.py` and `models/fields.py` -- the literal files, unmodified -- on top of a
pluggable (nerfacc, tinycudann) module pair.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Source of the code, in this order:
  * /root/reference/models/{renderer,fields}.py where the reference checkout exists (the build container);
  * oracle/_ref/models/{renderer,fields}.pyc, byte-compiled from those files by oracle/build_ref.py:build_models (the GPU
    box has no /root/reference; the compiled artefact travels with the snapshot like oracle/_ref/nerfacc_ref_C.so).

The two files import `mcubes`, `tqdm` (models/renderer.py:3-4), `icecream` (models/fields.py:5), `nerfacc`
(models/renderer.py:5-7) and `tinycudann` (models/fields.py:4).  `load(nerfacc_module, tcnn_module)` puts the given pair and
inert stubs for the three helpers into sys.modules for the duration of the import, then restores sys.modules.

Two pairs are used by tests/test_reference_literal.py:
  * CPU: `cpu_nerfacc()` / `cpu_tcnn()` below -- the oracle's own operators behind the nerfacc / tcnn calling conventions --
    which pins oracle/torch_ops.py's restatement of renderer.py / fields.py against the literal reference code;
  * CUDA: supernormal_b200.nerfacc_api / tcnn_api -- the drop-in claim itself.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types
from enum import Enum

import torch

from . import build_ref
from . import torch_ops as T

_SRC_DIR = build_ref.REF_MODELS


def available() -> bool:
    return os.path.isdir(_SRC_DIR) or build_ref.build_models() is not None


def _stub_modules():
    mc = types.ModuleType("mcubes")

    def marching_cubes(u, thr):   # models/renderer.py:29 -- not on the tested path
        raise NotImplementedError("mcubes stub: extract_geometry is covered by supernormal_b200.mesh")
    mc.marching_cubes = marching_cubes
    ic = types.ModuleType("icecream")
    ic.ic = lambda *a, **k: None
    stubs = {"mcubes": mc, "icecream": ic}
    try:
        import tqdm  # noqa: F401  (installed in this image; the reference only wraps a loop with it)
    except ImportError:
        tq = types.ModuleType("tqdm")
        tq.tqdm = lambda it, *a, **k: it
        stubs["tqdm"] = tq
    return stubs


def _exec(name: str):
    src = os.path.join(_SRC_DIR, name + ".py")
    if os.path.exists(src):
        loader = importlib.machinery.SourceFileLoader(f"_snb_ref_models_{name}", src)
    else:
        pyc = os.path.join(build_ref.MODELS_OUT, name + ".pyc")
        if not os.path.exists(pyc):
            raise FileNotFoundError(f"neither {src} nor {pyc}: run oracle/build_ref.py where /root/reference exists")
        loader = importlib.machinery.SourcelessFileLoader(f"_snb_ref_models_{name}", pyc)
    spec = importlib.util.spec_from_loader(loader.name, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode, saved = True, sys.dont_write_bytecode   # never write __pycache__ into /root/reference
    try:
        loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = saved
    return mod


def load(nerfacc_module, tcnn_module):
    """-> (renderer module, fields module): the literal reference files bound to the given operator modules."""
    inject = dict(_stub_modules(), nerfacc=nerfacc_module, tinycudann=tcnn_module)
    saved = {k: sys.modules.get(k) for k in inject}
    sys.modules.update(inject)
    try:
        return _exec("renderer"), _exec("fields")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


# ---- the oracle's CPU operators behind the nerfacc 0.3.5 / tinycudann calling conventions ---------------------------
class _ContractionType(Enum):   # CS/pybind.cu:165-168, NA/contraction.py:12-62
    AABB = 0
    UN_BOUNDED_TANH = 1
    UN_BOUNDED_SPHERE = 2


class _CpuGrid(T.OccupancyGrid):
    """nerfacc.OccupancyGrid(roi_aabb, resolution, contraction_type).to(device) as models/renderer.py:48-51 builds it."""

    def __init__(self, roi_aabb, resolution=128, contraction_type=_ContractionType.AABB):
        super().__init__(roi_aabb, resolution, device="cpu")
        self.contraction_type = contraction_type

    def to(self, *a, **k):   # the literal renderer asks for "cuda"; this pair computes on the CPU
        return self


def cpu_nerfacc() -> types.ModuleType:
    m = types.ModuleType("nerfacc")
    m.ContractionType = _ContractionType
    m.OccupancyGrid = _CpuGrid

    def ray_marching(rays_o, rays_d, t_min=None, t_max=None, scene_aabb=None, grid=None, sigma_fn=None, alpha_fn=None,
                     early_stop_eps=1e-4, alpha_thre=0.0, near_plane=None, far_plane=None, render_step_size=1e-3,
                     stratified=False, cone_angle=0.0):
        """NA/ray_marching.py:14-222 (signature), on oracle.torch_ops.ray_marching; the stratified jitter is drawn with
        torch.rand_like exactly where the reference draws it (:157-158), so a test reproduces it from the seed."""
        assert sigma_fn is None and t_min is not None and t_max is not None and grid is not None
        jitter = torch.rand_like(t_min) if stratified else None
        return T.ray_marching(rays_o, rays_d, t_min, t_max, grid.roi_aabb, grid.binary, render_step_size, cone_angle, alpha_fn,
                              early_stop_eps=early_stop_eps, alpha_thre=alpha_thre, jitter=jitter)

    m.ray_marching = ray_marching
    m.render_weight_from_alpha_patch_based = T.render_weight_from_alpha_patch_based
    m.accumulate_along_rays_patch_based = T.accumulate_along_rays_patch_based
    m.render_weight_from_alpha = T.render_weight_from_alpha
    m.accumulate_along_rays = T.accumulate_along_rays
    return m


def cpu_tcnn() -> types.ModuleType:
    """`tinycudann.Encoding` over oracle.torch_ops.hashgrid_encode: flat fp32 `params`, fp16-faithful output in fp16."""
    import oracle as _o

    class Encoding(torch.nn.Module):
        def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
            super().__init__()
            cfg = dict(encoding_config)
            self.spec = _o.hashgrid_spec(**{k: v for k, v in cfg.items() if k != "otype"})
            self.n_input_dims = n_input_dims
            self.n_output_dims = self.spec.n_levels * 2
            g = torch.Generator().manual_seed(seed)
            self.params = torch.nn.Parameter((torch.rand(self.spec.n_params, generator=g) * 2 - 1) * 1e-4)

        def forward(self, x):
            return T.hashgrid_encode(x, self.params, self.spec, fp16=True).to(torch.float16)

    m = types.ModuleType("tinycudann")
    m.Encoding = Encoding
    return m
