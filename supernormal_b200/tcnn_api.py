"""tiny-cuda-nn-shaped `Encoding(HashGrid)` module (boundary B of SURVEY.md §8b) on libsnb200.

Mirrors what models/fields.py:26-27,37,78 uses from tinycudann's torch bindings (tiny-cuda-nn
@2ec562e, bindings/torch/tinycudann/modules.py -- not vendored in the reference tree):
  * nn.Module with ONE flat fp32 nn.Parameter named `params` (level-major, entry-major,
    feature-minor; U(-1e-4,1e-4)), `.n_output_dims`, `.n_input_dims`;
  * forward(x f32[N,3]) -> fp16 [N, n_levels*2] computed from the fp16 cast of `params` with fp16
    trilinear accumulation; any N, inputs outside [0,1] wrap like the reference;
  * autograd: dL/dparams, dL/dx, and the double backward through dL/dx that
    SDFNetwork.gradient(create_graph=True) needs (models/fields.py:107-119).
Differences kept deliberately: the fp16 copy of the table is refreshed only when `params` changed
(tensor version counter) instead of on every call; table gradients accumulate in fp32 (tcnn: fp16
atomics with loss scale 128); no batch padding is needed.
"""
from __future__ import annotations

from typing import Mapping, Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import call, ptr


class _EncodeBwd(torch.autograd.Function):
    """First backward as a differentiable op (the tcnn bindings' _module_function_backward)."""

    @staticmethod
    def forward(ctx, enc, dout, x, params, table):
        dout = dout.contiguous()
        n = x.shape[0]
        dy_f32 = int(dout.dtype == torch.float32)
        dx = dparams = None
        if x.requires_grad:
            dx = torch.empty_like(x)
            call("snb_hashgrid_bwd_input", n, ptr(x), ptr(dout), dy_f32, ptr(table), enc._meta_ref, enc._n_active(), ptr(dx))
        if params.requires_grad:
            dparams = torch.zeros_like(params)
            call("snb_hashgrid_bwd_table", n, ptr(x), ptr(dout), dy_f32, 1.0, enc._meta_ref, enc._n_active(), ptr(dparams))
        ctx.enc = enc
        ctx.save_for_backward(dout, x, params, table)
        return dx, dparams

    @staticmethod
    def backward(ctx, g_dx, g_dparams):
        # supported: d(dL_dx)/d(dout), d(dL_dx)/d(params), d(dL_dx)/d(x)   (same set as tcnn)
        dout, x, params, table = ctx.saved_tensors
        enc = ctx.enc
        if g_dx is None:
            return None, None, None, None, None
        n = x.shape[0]
        g_dx = g_dx.contiguous().float()
        need_dout, need_x, need_p = ctx.needs_input_grad[1], ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        d_dout = torch.empty((n, enc.n_output_dims), device=x.device) if need_dout else None
        dx2 = torch.empty_like(x) if need_x else None
        gp = torch.zeros_like(params) if need_p else None
        call("snb_hashgrid_bwd_bwd_input", n, ptr(x), ptr(g_dx), ptr(dout), int(dout.dtype == torch.float32), ptr(table),
             enc._meta_ref, enc._n_active(), ptr(gp), ptr(d_dout), ptr(dx2))
        if d_dout is not None:
            d_dout = d_dout.to(dout.dtype)
        return None, d_dout, dx2, gp, None


class _Encode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, x, params):
        table = enc._table_f16(params)
        n = x.shape[0]
        out = torch.empty((n, enc.n_output_dims), dtype=enc.dtype, device=x.device)
        call("snb_hashgrid_fwd", n, ptr(x), ptr(table), enc._meta_ref, enc._n_active(), ptr(out), int(enc.dtype == torch.float32))
        ctx.enc = enc
        ctx.save_for_backward(x, params, table)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, params, table = ctx.saved_tensors
        dx, dparams = _EncodeBwd.apply(ctx.enc, dout, x, params, table)
        return None, dx, dparams


class Encoding(torch.nn.Module):
    """tcnn.Encoding(n_input_dims, encoding_config, seed=1337, dtype=None) for otype == "HashGrid"."""

    def __init__(self, n_input_dims: int, encoding_config: Mapping, seed: int = 1337, dtype: Optional[torch.dtype] = None):
        super().__init__()
        cfg = dict(encoding_config)
        otype = str(cfg.get("otype", "HashGrid"))
        if otype.lower() not in ("hashgrid", "grid"):
            raise NotImplementedError(f"supernormal_b200 implements otype=HashGrid only, got {otype}")
        if n_input_dims != 3:
            raise NotImplementedError("HashGrid with n_input_dims == 3 only (models/fields.py:26)")
        self.n_input_dims = n_input_dims
        self.n_levels = int(cfg.get("n_levels", 16))
        self.n_features_per_level = int(cfg.get("n_features_per_level", 2))
        if self.n_features_per_level != 2:
            raise NotImplementedError("n_features_per_level == 2 only (config/diligent.conf:82)")
        self.log2_hashmap_size = int(cfg.get("log2_hashmap_size", 19))
        self.base_resolution = int(cfg.get("base_resolution", 16))
        self.per_level_scale = float(cfg.get("per_level_scale", 2.0))
        self.encoding_config = cfg
        self.seed = seed
        self.dtype = torch.float16 if dtype is None else dtype
        if self.dtype not in (torch.float16, torch.float32):
            raise ValueError("dtype must be torch.float16 or torch.float32")
        self._meta, self.n_entries = _lib.make_meta(self.n_levels, self.log2_hashmap_size, self.base_resolution, self.per_level_scale)
        import ctypes
        self._meta_ref = ctypes.byref(self._meta)
        self.n_output_dims = self.n_levels * self.n_features_per_level
        g = torch.Generator().manual_seed(seed)
        init = (torch.rand(self.n_entries * 2, generator=g) * 2.0 - 1.0) * 1e-4
        self.params = torch.nn.Parameter(init)
        self.loss_scale = 1.0  # fp32 table gradients: no loss scaling needed (tcnn: 128 for fp16)
        self.n_active_levels: Optional[int] = None  # extension: skip levels >= this (outputs exact zeros)
        self._cache_key = None
        self._cache_table = None

    # -- helpers ----------------------------------------------------------------------------
    def _n_active(self) -> int:
        return self.n_levels if self.n_active_levels is None else max(0, min(int(self.n_active_levels), self.n_levels))

    def level_offsets(self):
        return [int(self._meta.offsets[i]) for i in range(self.n_levels + 1)]

    @torch.no_grad()
    def _table_f16(self, params: Tensor) -> Tensor:
        key = (params.data_ptr(), params._version, params.device)
        if key != self._cache_key:
            if self._cache_table is None or self._cache_table.device != params.device:
                self._cache_table = torch.empty(params.numel(), dtype=torch.float16, device=params.device)
            call("snb_cast_f32_to_f16", params.numel(), ptr(params.detach()), ptr(self._cache_table))
            self._cache_key = key
        return self._cache_table

    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")
        x = x.to(torch.float).contiguous()
        params = self.params if self.params.dtype == torch.float32 else self.params.float()
        return _Encode.apply(self, x, params.contiguous())

    def extra_repr(self):
        return f"n_input_dims={self.n_input_dims}, n_output_dims={self.n_output_dims}, seed={self.seed}, dtype={self.dtype}, hyperparams={self.encoding_config}"


# ---- tcnn.Network / tcnn.NetworkWithInputEncoding (FullyFusedMLP / CutlassMLP configs) ------------------------------------------------
# North-star API nicety ("optional"): SuperNormal itself never builds a tcnn network -- its SDF MLP is nn.Linear +
# weight_norm + Softplus(beta=100) in fp32 (models/fields.py:37-70) and runs in the fused kernels of trainer.py.  These two modules exist so
# that code written against tiny-cuda-nn's torch bindings (nerfacc's examples/radiance_fields/ngp.py:108-145) imports and runs on this
# package.  They follow the bindings' conventions -- one flat fp32 `params`, no biases, input / output widths padded to multiples of 16,
# Xavier-uniform init from `seed`, fp16 compute and fp16 output, `loss_scale` -- and evaluate the dense layers with torch's half-precision
# GEMMs (cuBLAS: a plain library GEMM, not a hot path of this repo).
_ACT = {
    "none": lambda x: x, "relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh, "exponential": torch.exp,
    "softplus": lambda x: torch.nn.functional.softplus(x * 10.0) / 10.0,          # tcnn: K_ACT = 10
    "squareplus": lambda x: 0.5 * (x * 10.0 + torch.sqrt(x * x * 100.0 + 4.0)) / 10.0,
    "sine": torch.sin, "leakyrelu": lambda x: torch.nn.functional.leaky_relu(x, 0.01),
}


def _pad16(n: int) -> int:
    return (n + 15) // 16 * 16


class Network(torch.nn.Module):
    """tcnn.Network(n_input_dims, n_output_dims, network_config, seed=1337): a bias-free MLP with `n_hidden_layers` hidden layers of
    `n_neurons` units.  params = [W_in (n_neurons x pad16(n_in)) | W_hidden ... | W_out (pad16(n_out) x n_neurons)], row-major."""

    def __init__(self, n_input_dims: int, n_output_dims: int, network_config: Mapping, seed: int = 1337):
        super().__init__()
        cfg = dict(network_config)
        otype = str(cfg.get("otype", "FullyFusedMLP")).lower()
        if otype not in ("fullyfusedmlp", "cutlassmlp", "megakernelmlp"):
            raise NotImplementedError(f"network otype {cfg.get('otype')!r}")
        self.n_input_dims, self.n_output_dims = int(n_input_dims), int(n_output_dims)
        self.n_neurons, self.n_hidden_layers = int(cfg.get("n_neurons", 64)), int(cfg.get("n_hidden_layers", 1))
        if self.n_hidden_layers < 1:
            raise ValueError("n_hidden_layers must be >= 1")
        self.activation, self.output_activation = str(cfg.get("activation", "ReLU")).lower(), str(cfg.get("output_activation", "None")).lower()
        for a in (self.activation, self.output_activation):
            if a not in _ACT:
                raise NotImplementedError(f"activation {a!r}")
        self.network_config, self.seed = cfg, seed
        self.padded_input, self.padded_output = _pad16(self.n_input_dims), _pad16(self.n_output_dims)
        self._shapes = [(self.n_neurons, self.padded_input)] + [(self.n_neurons, self.n_neurons)] * (self.n_hidden_layers - 1) + \
                       [(self.padded_output, self.n_neurons)]
        g = torch.Generator().manual_seed(seed)
        chunks = []
        for fo, fi in self._shapes:      # xavier_uniform over the padded matrix, like tcnn's initialize_params
            bound = (6.0 / (fi + fo)) ** 0.5
            chunks.append((torch.rand(fo * fi, generator=g) * 2.0 - 1.0) * bound)
        self.params = torch.nn.Parameter(torch.cat(chunks))
        self.loss_scale = 128.0
        self.dtype = torch.float16

    def _weights(self, params: Tensor):
        out, off = [], 0
        for fo, fi in self._shapes:
            out.append(params[off:off + fo * fi].view(fo, fi))
            off += fo * fi
        return out

    def mlp(self, x: Tensor, params: Tensor) -> Tensor:
        """x [N, n_input_dims] (any float dtype) -> fp16 [N, n_output_dims]"""
        h = x.to(torch.float16)
        if self.padded_input != self.n_input_dims:
            h = torch.nn.functional.pad(h, (0, self.padded_input - self.n_input_dims), value=1.0)   # tcnn pads inputs with ones
        ws = self._weights(params)
        for w in ws[:-1]:
            h = _ACT[self.activation](torch.nn.functional.linear(h, w.to(torch.float16)))
        y = _ACT[self.output_activation](torch.nn.functional.linear(h, ws[-1].to(torch.float16)))
        return y[:, :self.n_output_dims]

    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")
        return self.mlp(x, self.params)

    def extra_repr(self):
        return f"n_input_dims={self.n_input_dims}, n_output_dims={self.n_output_dims}, seed={self.seed}, dtype={self.dtype}, hyperparams={self.network_config}"


class NetworkWithInputEncoding(torch.nn.Module):
    """tcnn.NetworkWithInputEncoding(n_input_dims, n_output_dims, encoding_config, network_config, seed=1337): HashGrid encoding ->
    Network, ONE flat parameter vector params = [network | encoding] (the order of tiny-cuda-nn's NetworkWithInputEncoding::set_params)."""

    def __init__(self, n_input_dims: int, n_output_dims: int, encoding_config: Mapping, network_config: Mapping, seed: int = 1337):
        super().__init__()
        enc = Encoding(n_input_dims, encoding_config, seed=seed)
        net = Network(enc.n_output_dims, n_output_dims, network_config, seed=seed)
        self.n_input_dims, self.n_output_dims, self.seed = int(n_input_dims), int(n_output_dims), seed
        self._n_net = net.params.numel()
        self.params = torch.nn.Parameter(torch.cat([net.params.detach(), enc.params.detach()]))
        del enc._parameters["params"], net._parameters["params"]     # the flat vector above is the only parameter (checkpoints, optimizers)
        object.__setattr__(self, "_enc", enc)                        # helper objects, deliberately not registered sub-modules
        object.__setattr__(self, "_net", net)
        self.loss_scale, self.dtype = 128.0, torch.float16

    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")
        p_net, p_enc = self.params[:self._n_net], self.params[self._n_net:]
        feat = _Encode.apply(self._enc, x.to(torch.float).contiguous(), p_enc.contiguous())
        return self._net.mlp(feat, p_net)


def install_as_tinycudann():
    """Register this module as `tinycudann` (and nerfacc_api as `nerfacc`) in sys.modules so the
    reference's models/fields.py / models/renderer.py import unmodified."""
    import sys
    from . import nerfacc_api
    sys.modules.setdefault("tinycudann", sys.modules[__name__])
    sys.modules.setdefault("nerfacc", nerfacc_api)
