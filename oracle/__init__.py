"""oracle -- CPU restatement of the reference algorithm on SuperNormal's patch-based NeuS
hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product (supernormal_b200/) never does and fails loudly when
its CUDA library is missing instead of falling back to anything here.

Layout
  oracle/c/oracle.c     bit-faithful C restatement of the reference CUDA kernels (ray
                        marching, patch weights fwd/bwd, serial transmittance) and of
                        tiny-cuda-nn's hash-grid kernels (fp16-faithful forward).
  oracle/torch_ops.py   PyTorch-CPU expression of the whole training step (nerfacc Python
                        wrappers, models/renderer.py, models/fields.py, exp_runner.py loss +
                        Adam), differentiable to 2nd order.  Also the reported CPU baseline.
  oracle/build_ref.py   recipe that compiles the UNMODIFIED reference nerfacc CUDA extension
                        into oracle/_ref/ for on-GPU differential tests.

Parity status: nerfacc operators are pinned by the reference's golden vectors
(tests/test_oracle_golden.py) and by oracle/_ref on the GPU box.  The hash grid
(tiny-cuda-nn, absent from the reference tree, pin create_env.sh:11) and marching cubes
(PyMCubes 0.1.4, absent) are PARITY UNPINNED: restated from their published algorithms.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    """Compile oracle/c/oracle.c with gcc (seconds). Returns the .so path."""
    out = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "c", "oracle.c")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return out


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_ray_marching.restype = C.c_int64
        _LIB.oracle_hashgrid_offsets.restype = C.c_uint32
        _LIB.oracle_grid_scale.restype = C.c_float
        _LIB.oracle_grid_scale.argtypes = [C.c_uint32, C.c_float, C.c_uint32]
        _LIB.oracle_f32_to_h.restype = C.c_uint16
        _LIB.oracle_f32_to_h.argtypes = [C.c_float]
        _LIB.oracle_h_to_f32.restype = C.c_float
        _LIB.oracle_h_to_f32.argtypes = [C.c_uint16]
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


# ----------------------------------------------------------------------------------------
# nerfacc kernels
# ----------------------------------------------------------------------------------------
def ray_marching(rays_o, rays_d, t_min, t_max, roi, grid_binary, step_size, cone_angle=0.0):
    """CS/ray_marching.cu:194-289. Returns (packed_info i32[n,2], ray_indices i64[S],
    t_starts f32[S,1], t_ends f32[S,1])."""
    rays_o, rays_d, t_min, t_max, roi = map(_f32, (rays_o, rays_d, t_min, t_max, roi))
    grid = np.ascontiguousarray(np.asarray(grid_binary).astype(np.uint8))
    assert grid.ndim == 3 and rays_o.shape == rays_d.shape and rays_o.shape[1] == 3
    n = rays_o.shape[0]
    res = np.asarray(grid.shape, dtype=np.int32)
    packed = np.zeros((n, 2), dtype=np.int32)
    args = (C.c_int(n), _p(rays_o), _p(rays_d), _p(t_min), _p(t_max), _p(roi), _p(res), _p(grid),
            C.c_float(step_size), C.c_float(cone_angle), _p(packed))
    total = lib().oracle_ray_marching(*args, None, None, None)
    ridx = np.zeros(total, dtype=np.int64)
    t0 = np.zeros((total, 1), dtype=np.float32)
    t1 = np.zeros((total, 1), dtype=np.float32)
    lib().oracle_ray_marching(*args, _p(ridx), _p(t0), _p(t1))
    return packed, ridx, t0, t1


def weight_from_alpha_patch_fwd(packed_info, alphas):
    """CS/render_weight.cu:87-118. alphas f32[S,P,1] (or [S,1] for the per-ray kernel)."""
    alphas = _f32(alphas)
    packed_info = np.ascontiguousarray(packed_info, dtype=np.int32)
    P = alphas.shape[1] if alphas.ndim == 3 else 1
    w = np.zeros_like(alphas)
    lib().oracle_weight_from_alpha_patch_fwd(C.c_int(packed_info.shape[0]), C.c_int(P), _p(packed_info),
                                             _p(alphas), _p(w))
    return w


def weight_from_alpha_patch_bwd(packed_info, alphas, weights, grad_weights):
    """CS/render_weight.cu:298-340."""
    alphas, weights, grad_weights = map(_f32, (alphas, weights, grad_weights))
    packed_info = np.ascontiguousarray(packed_info, dtype=np.int32)
    P = alphas.shape[1] if alphas.ndim == 3 else 1
    g = np.zeros_like(alphas)
    lib().oracle_weight_from_alpha_patch_bwd(C.c_int(packed_info.shape[0]), C.c_int(P), _p(packed_info),
                                             _p(alphas), _p(weights), _p(grad_weights), _p(g))
    return g


def transmittance_from_alpha(packed_info, alphas):
    """CS/render_transmittance.cu:85-112 (serial order)."""
    alphas = _f32(alphas)
    packed_info = np.ascontiguousarray(packed_info, dtype=np.int32)
    T = np.zeros_like(alphas)
    lib().oracle_transmittance_from_alpha(C.c_int(packed_info.shape[0]), _p(packed_info), _p(alphas), _p(T))
    return T


# ----------------------------------------------------------------------------------------
# tiny-cuda-nn hash grid (parity unpinned: source absent from the reference tree)
# ----------------------------------------------------------------------------------------
@dataclass
class HashGridSpec:
    n_levels: int
    n_features: int
    log2_hashmap_size: int
    base_resolution: int
    per_level_scale: float
    offsets: np.ndarray      # uint32 [L+1], in entries
    scales: np.ndarray       # float32 [L]
    resolutions: np.ndarray  # uint32 [L]

    @property
    def n_entries(self) -> int:
        return int(self.offsets[-1])

    @property
    def n_params(self) -> int:
        return self.n_entries * self.n_features

    @property
    def n_output_dims(self) -> int:
        return self.n_levels * self.n_features


def hashgrid_spec(n_levels=14, n_features_per_level=2, log2_hashmap_size=19, base_resolution=32,
                  per_level_scale=1.3195079107728942, **_ignored) -> HashGridSpec:
    """Offset table of GridEncodingTemplated (tcnn encodings/grid.h), SURVEY Appendix A.1-2."""
    assert n_features_per_level == 2, "the hot path uses F=2 (config/diligent.conf:80-87)"
    off = np.zeros(n_levels + 1, dtype=np.uint32)
    sc = np.zeros(n_levels, dtype=np.float32)
    rs = np.zeros(n_levels, dtype=np.uint32)
    lib().oracle_hashgrid_offsets(C.c_uint32(n_levels), C.c_uint32(log2_hashmap_size),
                                  C.c_uint32(base_resolution), C.c_float(per_level_scale),
                                  _p(off), _p(sc), _p(rs))
    return HashGridSpec(n_levels, 2, log2_hashmap_size, base_resolution, per_level_scale, off, sc, rs)


def hashgrid_fwd(spec: HashGridSpec, x, table_f16, n_active=None, want_dy_dx=False):
    """tcnn kernel_grid forward, fp16-faithful. x f32[N,3]; table_f16 np.float16[n_entries*2].
    Returns np.float16 [N, L*2] (and dy_dx f32 [N, L*2, 3])."""
    x = _f32(x)
    t = np.ascontiguousarray(np.asarray(table_f16, dtype=np.float16)).view(np.uint16)
    assert t.size == spec.n_params
    n = x.shape[0]
    na = spec.n_levels if n_active is None else int(n_active)
    out = np.zeros((n, spec.n_output_dims), dtype=np.uint16)
    dydx = np.zeros((n, spec.n_output_dims, 3), dtype=np.float32) if want_dy_dx else None
    lib().oracle_hashgrid_fwd(C.c_int64(n), _p(x), _p(t), C.c_uint32(spec.n_levels), _p(spec.offsets),
                              _p(spec.scales), _p(spec.resolutions), C.c_uint32(na), _p(out),
                              _p(dydx) if want_dy_dx else None)
    out = out.view(np.float16)
    return (out, dydx) if want_dy_dx else out


def hashgrid_bwd_table(spec: HashGridSpec, x, dL_dy, n_active=None):
    """tcnn kernel_grid_backward with exact (double) accumulation. Returns f64 [n_entries*2]."""
    x, dL_dy = _f32(x), _f32(dL_dy)
    na = spec.n_levels if n_active is None else int(n_active)
    g = np.zeros(spec.n_params, dtype=np.float64)
    lib().oracle_hashgrid_bwd_table(C.c_int64(x.shape[0]), _p(x), _p(dL_dy), C.c_uint32(spec.n_levels),
                                    _p(spec.offsets), _p(spec.scales), _p(spec.resolutions),
                                    C.c_uint32(na), _p(g))
    return g


def hashgrid_corners(spec: HashGridSpec, x):
    """Absolute entry index (int64 [N,L,8]) and trilinear weight (f32 [N,L,8]) of each corner."""
    x = _f32(x)
    n = x.shape[0]
    idx = np.zeros((n, spec.n_levels, 8), dtype=np.int64)
    w = np.zeros((n, spec.n_levels, 8), dtype=np.float32)
    lib().oracle_hashgrid_corners(C.c_int64(n), _p(x), C.c_uint32(spec.n_levels), _p(spec.offsets),
                                  _p(spec.scales), _p(spec.resolutions), _p(idx), _p(w))
    return idx, w
