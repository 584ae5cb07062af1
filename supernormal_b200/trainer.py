"""Fused, sync-free training step for SuperNormal's patch-based NeuS (the hot loop of
exp_runner.py:147-210 = Runner.train, around models/renderer.py:63-276 and models/fields.py:76-99).

Host side only: schedules, buffer ownership and ~10 C-ABI calls per iteration.  All tensor math runs
in libsnb200 (hand-written sm_100a kernels); there is no PyTorch/CPU fallback.  State-dict keys of the
reference (`encoding.params`, `lin{0,1}.{weight_g,weight_v,bias}`, `variance`) are
produced / consumed by `reference_state_dict` / `load_reference_state_dict`.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, dp
from ._lib import HashGridMeta, call, ptr

H = 64
NET_FLOATS = 2432
P = 9
SMALL_PAD = 2560  # floats reserved for the MLP/variance parameters in front of the hash table (16B aligned)


class SnbNet(C.Structure):
    _fields_ = [("table_f16", C.c_void_p), ("net", C.c_void_p), ("meta", HashGridMeta), ("n_active", C.c_uint32)]


class SnbPatchBatch(C.Structure):
    _fields_ = [("n_patches", C.c_int32), ("rays_o", C.c_void_p), ("rays_d", C.c_void_p), ("plane_n", C.c_void_p),
                ("near_", C.c_void_p), ("far_", C.c_void_p), ("v_inv", C.c_void_p), ("normal_gt", C.c_void_p),
                ("mask", C.c_void_p)]


class SnbSamples(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("end_capacity", C.c_int64), ("scratch_stride", C.c_int32),
                ("counts", C.c_void_p), ("end_counts", C.c_void_p), ("packed_info", C.c_void_p),
                ("end_packed", C.c_void_p), ("totals", C.c_void_p), ("t0", C.c_void_p), ("t1", C.c_void_p),
                ("patch_idx", C.c_void_p), ("end_slot", C.c_void_p), ("slot_sample", C.c_void_p),
                ("scratch_t0", C.c_void_p), ("scratch_t1", C.c_void_p), ("launch_order", C.c_void_p)]


class SnbDataset(C.Structure):
    _fields_ = [("n_images", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("n_train", C.c_int32), ("normals", C.c_void_p),
                ("masks", C.c_void_p), ("intrinsics_inv", C.c_void_p), ("pose", C.c_void_p), ("v_inverse", C.c_void_p),
                ("train_ids", C.c_void_p)]


class SnbBatchOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rays_o", "rays_d", "plane_n", "near_", "far_", "v_inv", "normal_gt", "mask", "jitter")]


class SnbTrainCtx(C.Structure):
    _fields_ = [("batch", SnbPatchBatch), ("samples", SnbSamples), ("net", SnbNet), ("n_levels", C.c_int32), ("small_pad", C.c_int64),
                ("flat_param", C.c_void_p), ("flat_grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("net_grad", C.c_void_p), ("sdf", C.c_void_p), ("feats", C.c_void_p), ("d_sdf0", C.c_void_p), ("d_sdf1", C.c_void_p),
                ("comp", C.c_void_p), ("wsum", C.c_void_p), ("dcomp", C.c_void_p), ("dwsum", C.c_void_p), ("stats", C.c_void_p),
                ("jitter", C.c_void_p), ("roi", C.c_void_p), ("grid_binary", C.c_void_p), ("res_x", C.c_int32), ("res_y", C.c_int32),
                ("res_z", C.c_int32), ("bwd_workspace", C.c_void_p), ("bwd_workspace_bytes", C.c_int64)]


class SnbPeerGroup(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("param", C.c_void_p * dp.MAX_PEERS), ("grad", C.c_void_p * dp.MAX_PEERS),
                ("table_f16", C.c_void_p * dp.MAX_PEERS), ("flags", C.c_void_p * dp.MAX_PEERS), ("counter", C.c_void_p),
                ("table_f16_only", C.c_int32), ("epoch", C.c_uint32), ("trace", C.c_void_p), ("trace_capacity", C.c_int32)]


class SDFModel:
    """Parameters of SDFNetwork (models/fields.py:7-99, shipped configuration: one 64-wide hidden layer,
    weight-norm, Softplus(beta=100), input_concat, geometric init) + SingleVarianceNetwork
    (models/fields.py:133-139) in ONE flat fp32 buffer:  [ small | pad | hash table ]
        small = v0[64, d_in] | g0[64] | b0[64] | v1[64] | g1 | b1 | variance,   d_in = 3 + 2*n_levels
    so that Adam and the data-parallel gradient exchange are single contiguous sweeps.  A persistent fp16 copy
    of the table (what tiny-cuda-nn gathers from) is refreshed by the optimizer kernel."""

    def __init__(self, encoding_config: dict, bias: float = 0.6, variance_init: float = 0.5, seed: int = 1337, device="cuda",
                 peer_group_factory=None):
        """peer_group_factory(nbytes) -> dp.PeerGroup: put parameters, gradients and the fp16 table into a symmetric (peer-mapped)
        allocation so that the data-parallel tail kernel (snb_train_tail_peer) can reduce / broadcast them over NVLink."""
        cfg = dict(encoding_config)
        self.n_levels = int(cfg["n_levels"])
        assert int(cfg.get("n_features_per_level", 2)) == 2
        self.meta, self.n_entries = _lib.make_meta(self.n_levels, int(cfg["log2_hashmap_size"]), int(cfg["base_resolution"]),
                                                   float(cfg["per_level_scale"]))
        self.offsets = [int(self.meta.offsets[i]) for i in range(self.n_levels + 1)]
        self.d_in = 3 + 2 * self.n_levels
        self.n_small = H * self.d_in + 3 * H + 3
        assert self.n_small <= SMALL_PAD
        self.n_table = self.n_entries * 2
        self.device = torch.device(device)
        g = torch.Generator().manual_seed(seed)
        flat = torch.zeros(SMALL_PAD + self.n_table)
        flat[SMALL_PAD:] = (torch.rand(self.n_table, generator=g) * 2 - 1) * 1e-4   # tcnn init U(-1e-4, 1e-4)
        # geometric init, models/fields.py:47-65 (l == 0 with encoding; l == num_layers-2), then weight_norm: g = |v|
        w0 = torch.zeros(H, self.d_in)
        w0[:, :3].normal_(0.0, math.sqrt(2) / math.sqrt(H), generator=g)
        w1 = torch.empty(H).normal_(math.sqrt(math.pi) / math.sqrt(H), 1e-4, generator=g)
        o = self._small_offsets()
        flat[o["v0"]:o["g0"]] = w0.flatten()
        flat[o["g0"]:o["b0"]] = w0.norm(dim=1)
        flat[o["v1"]:o["g1"]] = w1
        flat[o["g1"]] = w1.norm()
        flat[o["b1"]] = -bias
        flat[o["var"]] = variance_init
        self.peer = None
        if peer_group_factory is not None:
            n_flat = SMALL_PAD + self.n_table
            offs, total = dp.carve_layout([n_flat * 4, n_flat * 4, self.n_table * 2, dp.PEER_FLAG_WORDS * 4, 256])
            pg = peer_group_factory(total)            # zero-filled, all ranks synchronised
            self.peer, self.peer_offsets = pg, offs
            self.flat = pg.view(offs[0], n_flat, torch.float32)
            self.flat.copy_(flat)
            self.grad = pg.view(offs[1], n_flat, torch.float32)
            self.table_f16 = pg.view(offs[2], self.n_table, torch.float16)
            self.peer_flags = pg.view(offs[3], dp.PEER_FLAG_WORDS, torch.int32)
            self.peer_counter = pg.view(offs[4], 1, torch.int32)
            self.net_grad = self.grad[:NET_FLOATS]    # the backward accumulates the folded-weight gradient where the peers read it
            # owners broadcast only the fp16 table copy the gathers read (2 instead of 14 bytes per owned parameter leave the GPU); fp32 table
            # parameters are then valid on the owner's chunks only and gather_table() reassembles them.  Validated on 2 GPUs in round 2
            # (profiles/r02_dp_peer_check_n2_f16only*.json: replicas bit-identical, trajectory equal to the NCCL path, full schedule 2.29 ->
            # 2.16 s); SNB_PEER_F16ONLY=0 keeps the fp32-broadcast form (the one validated at N=8 in round 1) as the cross-check.
            self.peer_f16_only = os.environ.get("SNB_PEER_F16ONLY", "1") == "1"
        else:
            self.flat = flat.to(self.device)
            self.grad = torch.zeros_like(self.flat)
            self.table_f16 = torch.empty(self.n_table, dtype=torch.float16, device=self.device)
            self.net_grad = torch.zeros(NET_FLOATS, device=self.device)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.net = torch.zeros(NET_FLOATS, device=self.device)
        self.n_active = 0  # SDFNetwork.bindwidth
        self.net_stale = True  # `net` / `net_grad` are not (folded current parameters / zero): the lean step calls prep() first
        self.refresh_table_f16()

    def _small_offsets(self) -> Dict[str, int]:
        v0 = 0
        g0 = H * self.d_in
        b0 = g0 + H
        v1 = b0 + H
        g1 = v1 + H
        return dict(v0=v0, g0=g0, b0=b0, v1=v1, g1=g1, b1=g1 + 1, var=g1 + 2)

    @property
    def small(self):
        return self.flat[:self.n_small]

    @property
    def table(self):
        return self.flat[SMALL_PAD:]

    def refresh_table_f16(self):
        call("snb_cast_f32_to_f16", self.n_table, ptr(self.table), ptr(self.table_f16))

    def peer_struct(self) -> SnbPeerGroup:
        pg, o = self.peer, self.peer_offsets
        arr = lambda off: (C.c_void_p * dp.MAX_PEERS)(*(pg.ptrs(off) + [None] * (dp.MAX_PEERS - pg.world)))
        # SNB_PEER_TRACE=n: keep the in-kernel timeline (4 globaltimer stamps of block 0) of the last n peer-tail launches (scripts/dp_peer_trace.py)
        n_trace = int(os.environ.get("SNB_PEER_TRACE", "0"))
        self.peer_trace = torch.zeros(max(n_trace, 1), 4, dtype=torch.int64, device=self.device) if n_trace > 0 else None
        return SnbPeerGroup(pg.world, pg.rank, arr(o[0]), arr(o[1]), arr(o[2]), arr(o[3]), self.peer_counter.data_ptr(),
                            int(self.peer_f16_only), 0, None if self.peer_trace is None else self.peer_trace.data_ptr(), n_trace)

    @torch.no_grad()
    def gather_table(self) -> torch.Tensor:
        """The full fp32 table on every rank.  Collective when the peer tail keeps fp32 parameters on their owners only
        (SNB_PEER_F16ONLY=1): non-owned chunks are zeroed in a copy and the copies summed; otherwise a plain clone."""
        t = self.table.clone()
        if self.peer is not None and self.peer_f16_only:
            import torch.distributed as dist
            t *= dp.owner_mask(self.n_table, self.peer.rank, self.peer.world, self.device)
            dist.all_reduce(t)
        return t

    def net_struct(self) -> SnbNet:
        return SnbNet(self.table_f16.data_ptr(), self.net.data_ptr(), self.meta, self.n_active)

    def prep(self, mask: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None):
        """weight_norm + variance folding (-> self.net); optionally mask_sum and stats reset."""
        call("snb_prep_net", self.n_levels, ptr(self.small), ptr(self.net), 0 if mask is None else mask.numel(),
             ptr(mask), ptr(stats))

    @torch.no_grad()
    def sdf(self, x: torch.Tensor, mode: int = 0) -> torch.Tensor:
        """SDFNetwork.sdf under no_grad (mode 1: occ_eval_fn of models/renderer.py:56-60, mode 2: -sdf)."""
        x = x.contiguous().float()
        out = torch.empty(x.shape[0], device=x.device)
        net = self.net_struct()
        call("snb_sdf_eval", x.shape[0], ptr(x), C.byref(net), mode, ptr(out))
        return out.unsqueeze(-1)

    @torch.no_grad()
    def sdf_and_gradient(self, x: torch.Tensor):
        """(sdf [n,1], d sdf/dx [n,3]) in one pass -- SDFNetwork.gradient (models/fields.py:107-119) for `ad`
        normals, forward-mode on the tensor cores instead of an autograd double pass."""
        x = x.contiguous().float()
        sdf = torch.empty(x.shape[0], device=x.device)
        grad = torch.empty(x.shape[0], 3, device=x.device)
        net = self.net_struct()
        call("snb_sdf_eval_grad", x.shape[0], ptr(x), C.byref(net), ptr(sdf), ptr(grad))
        return sdf.unsqueeze(-1), grad

    # -- reference checkpoint format (exp_runner.py:298-315) --------------------------------------
    def reference_state_dict(self) -> Dict[str, torch.Tensor]:
        o, s = self._small_offsets(), self.small
        return {
            "sdf_network_fine": {
                "encoding.params": self.gather_table(),
                "lin0.bias": s[o["b0"]:o["v1"]].clone(), "lin0.weight_g": s[o["g0"]:o["b0"]].clone().view(H, 1),
                "lin0.weight_v": s[o["v0"]:o["g0"]].clone().view(H, self.d_in),
                "lin1.bias": s[o["b1"]:o["b1"] + 1].clone(), "lin1.weight_g": s[o["g1"]:o["g1"] + 1].clone().view(1, 1),
                "lin1.weight_v": s[o["v1"]:o["g1"]].clone().view(1, H)},
            "variance_network_fine": {"variance": s[o["var"]].clone()},
        }

    @torch.no_grad()
    def load_reference_state_dict(self, sd: Dict[str, Dict[str, torch.Tensor]]):
        o, s = self._small_offsets(), self.small
        f = {k: v.to(self.device).float() for k, v in sd["sdf_network_fine"].items()}
        self.table.copy_(f["encoding.params"].flatten())
        s[o["v0"]:o["g0"]] = f["lin0.weight_v"].flatten()
        s[o["g0"]:o["b0"]] = f["lin0.weight_g"].flatten()
        s[o["b0"]:o["v1"]] = f["lin0.bias"].flatten()
        s[o["v1"]:o["g1"]] = f["lin1.weight_v"].flatten()
        s[o["g1"]] = f["lin1.weight_g"].flatten()[0]
        s[o["b1"]] = f["lin1.bias"].flatten()[0]
        s[o["var"]] = sd["variance_network_fine"]["variance"].to(self.device).float()
        self.refresh_table_f16()
        self.net_stale = True


class SampleBuffers:
    """Capacity buffers for the data-dependent sample lists (no host sync, SURVEY §7 'Dynamic sizes')."""

    def __init__(self, n_patches: int, samples_per_ray_cap: int, scratch_stride: int, n_levels: int, device):
        self.n_patches = n_patches
        self.capacity = n_patches * samples_per_ray_cap
        self.end_capacity = max(n_patches * 8, self.capacity // 4)
        i32 = dict(dtype=torch.int32, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        self.counts = torch.zeros(n_patches, **i32)
        self.end_counts = torch.zeros(n_patches, **i32)
        self.packed_info = torch.zeros(n_patches, 2, **i32)
        self.end_packed = torch.zeros(n_patches, 2, **i32)
        # loss accumulators (8 f32) and sample totals (4 i32) share ONE 48-byte allocation: the host-fed path reads both back with one D2H
        # copy.  Two such sets: HostBatchFeeder alternates them step by step (flip()), so the read-back of step i can run on a side stream
        # while step i + 1 already resets / accumulates into the other set.  Everything else uses set 0 throughout.
        self._sets = [torch.zeros(12, **i32) for _ in range(2)]
        self.set_idx = 0
        self.t0 = torch.zeros(self.capacity, **f32)
        self.t1 = torch.zeros(self.capacity, **f32)
        self.patch_idx = torch.zeros(self.capacity, **i32)
        self.end_slot = torch.zeros(self.capacity, **i32)
        self.slot_sample = torch.zeros(self.end_capacity, **i32)
        self.scratch_stride = scratch_stride
        self.scratch_t0 = torch.zeros(n_patches * scratch_stride, **f32)
        self.scratch_t1 = torch.zeros(n_patches * scratch_stride, **f32)
        self.launch_order = torch.arange(n_patches, **i32)        # patches by sample count, longest first (written by the compaction)
        m_cap = P * (self.capacity + self.end_capacity)
        self.sdf = torch.zeros(m_cap, **f32)
        self.feats = torch.zeros(m_cap * 16 * 2, dtype=torch.float16, device=device)   # rows of feat_row_stride(n_active) <= 16 half2 (fused_sdf.cu)
        self.d_sdf0 = torch.zeros(P * self.capacity, **f32)
        self.d_sdf1 = torch.zeros(P * self.capacity, **f32)
        self.comp = torch.zeros(n_patches, P, 3, **f32)
        self.wsum = torch.zeros(n_patches, P, **f32)
        self.dcomp = torch.zeros(n_patches, P, 3, **f32)
        self.dwsum = torch.zeros(n_patches, P, **f32)
        # workspace of the split backward (snb_sdf_bwd_patch_ws): positions + d loss / d features of every point, [level][ray][sample]
        self.bwd_ws_bytes = int(_lib.lib().snb_sdf_bwd_workspace_bytes(n_levels, self.capacity, self.end_capacity))
        self.bwd_ws = torch.empty(self.bwd_ws_bytes, dtype=torch.uint8, device=device)
        self._structs = [SnbSamples(self.capacity, self.end_capacity, scratch_stride, *[t.data_ptr() for t in (
            self.counts, self.end_counts, self.packed_info, self.end_packed, st[8:12], self.t0, self.t1,
            self.patch_idx, self.end_slot, self.slot_sample, self.scratch_t0, self.scratch_t1, self.launch_order)]) for st in self._sets]

    # the current accumulator set (see __init__)
    @property
    def stats_totals(self) -> torch.Tensor:
        return self._sets[self.set_idx]

    @property
    def stats(self) -> torch.Tensor:
        return self._sets[self.set_idx][:8].view(torch.float32)

    @property
    def totals(self) -> torch.Tensor:
        return self._sets[self.set_idx][8:12]

    @property
    def struct(self) -> SnbSamples:
        return self._structs[self.set_idx]

    def flip(self) -> None:
        self.set_idx ^= 1


BATCH_KEYS = ("rays_o", "rays_d", "plane_n", "near", "far", "v_inv", "normal_gt", "mask")


class HostBatchFeeder:
    """Feeds FusedTrainer.train_step from HOST patch batches (what Dataset.gen_random_patches would produce on a CPU
    loader, models/dataset_loader.py:223-277) and returns the loss terms to the host, without stalling the GPU:
      * every batch (8 tensors + the stratified jitter) is packed into ONE pinned staging buffer and moved with ONE
        cudaMemcpyAsync on a copy stream into one of `depth` device slots, overlapping the previous step's kernels;
      * the compute stream waits on the slot's event only; the slot is released by an event after the step;
      * the step's loss terms (stats[8], totals[4]) are copied device->host into a pinned ring every step (async);
        `losses()` synchronises once and decodes them."""

    def __init__(self, tr: "FusedTrainer", depth: int = 2, log_capacity: int = 4096):
        self.tr, self.depth = tr, depth
        n = tr.n_patches
        shapes = dict(rays_o=(n, 3), rays_d=(n, P, 3), plane_n=(n, 3), near=(n,), far=(n,), v_inv=(n, P, 9), normal_gt=(n, P, 3),
                      mask=(n, P), jitter=(n,))
        self.layout, off = {}, 0
        for k, shp in shapes.items():
            cnt = int(np.prod(shp))
            self.layout[k] = (off, cnt, shp)
            off += (cnt + 3) // 4 * 4          # keep every field 16-byte aligned
        self.floats = off
        self.h2d_bytes = off * 4
        self.d2h_bytes = 12 * 4
        self.host = [torch.empty(off, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.dev = [torch.empty(off, dtype=torch.float32, device=tr.device) for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=tr.device)
        self.d2h_stream = torch.cuda.Stream(device=tr.device)      # loss read-backs: never in front of a batch copy, never inside the compute stream
        self.step_done = [torch.cuda.Event() for _ in range(2)]
        self.d2h_done = [torch.cuda.Event() for _ in range(2)]
        self.log = torch.zeros(log_capacity, 12, dtype=torch.int32).pin_memory()      # per step: stats[8] (f32 bits) | totals[4]
        self.n_fed = self.n_run = 0
        # per-step host work is kept to a handful of driver calls (scripts/host_profile_e2e.py: at a 0.19 ms device step the interpreter
        # is the next bottleneck): field views of every device slot are built once, the copies go through snb_copy_async with cached
        # raw pointers / stream handles instead of Tensor.copy_ under a stream context manager, and the compute stream object is looked
        # up again only when torch's current raw stream changes
        self._slot_batch = [{k_: v for k_, v in self._views(d).items() if k_ != "jitter"} for d in self.dev]
        self._slot_jitter = [self._views(d)["jitter"] for d in self.dev]
        self._dev_ptr = [d.data_ptr() for d in self.dev]
        self._log_ptr, self._log_row_bytes = self.log.data_ptr(), 12 * 4
        self._set_ptr = [st.data_ptr() for st in tr.buf._sets]
        self._copy_raw, self._d2h_raw = self.copy_stream.cuda_stream, self.d2h_stream.cuda_stream
        self._copy_async = _lib.lib().snb_copy_async
        self._cur, self._cur_raw = None, None
        self._src_keep = [None] * depth

    def _compute_stream(self) -> "torch.cuda.Stream":
        raw = _lib.stream()
        if raw != self._cur_raw:
            self._cur, self._cur_raw = torch.cuda.current_stream(self.tr.device), raw
        return self._cur

    def _views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {k: flat[o:o + c].view(shp) for k, (o, c, shp) in self.layout.items()}

    def new_host_batch(self) -> torch.Tensor:
        """A pinned, packed host batch for a loader to fill through `views()` (zero-copy hand-over to submit())."""
        return torch.empty(self.floats, dtype=torch.float32).pin_memory()

    def views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Named field views (rays_o, rays_d, ..., mask, jitter) of a packed batch buffer."""
        return self._views(flat)

    def pack(self, batch: Dict[str, torch.Tensor], jitter: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Copy an unpacked host batch into a packed pinned buffer (what a loader does when it cannot fill views() directly)."""
        out = self.new_host_batch() if out is None else out
        hv = self._views(out)
        for k in BATCH_KEYS:
            hv[k].copy_(batch[k].reshape(hv[k].shape))
        hv["jitter"].copy_(jitter)
        return out

    def submit(self, batch, jitter: Optional[torch.Tensor] = None) -> None:
        """Start the H2D copy of one host batch into the next device slot (returns immediately).  `batch` is a packed
        pinned buffer from new_host_batch()/pack() -- one cudaMemcpyAsync -- or an unpacked dict (+ jitter), packed here."""
        if self.n_fed - self.n_run >= self.depth:
            raise RuntimeError(f"HostBatchFeeder.submit: {self.n_fed - self.n_run} batches already queued (depth {self.depth}); "
                               "call step() before submitting more -- the oldest slot has not been consumed")
        slot = self.n_fed % self.depth
        if self.n_fed >= self.depth:
            self.free[slot].synchronize()            # the step that used this slot `depth` submissions ago has finished
        src = batch if torch.is_tensor(batch) else self.pack(batch, jitter, self.host[slot])
        assert src.is_pinned() and src.numel() == self.floats
        self._src_keep[slot] = src                     # the copy is asynchronous: the packed host buffer must outlive it (until the slot's next use)
        _lib.check(self._copy_async(self._dev_ptr[slot], src.data_ptr(), self.h2d_bytes, 1, self._copy_raw))
        self.ready[slot].record(self.copy_stream)
        self.n_fed += 1

    def step(self) -> None:
        """Run one training iteration on the oldest submitted batch; queue the async read-back of its loss terms."""
        assert self.n_run < self.n_fed, "submit() a batch first"
        tr, slot = self.tr, self.n_run % self.depth
        cur = self._compute_stream()
        cur.wait_event(self.ready[slot])
        tr.buf.flip()                                   # this step accumulates into the set the step before last used ...
        k = tr.buf.set_idx
        cur.wait_event(self.d2h_done[k])                # ... whose read-back (two steps ago) must have left the device
        tr.train_step(batch=self._slot_batch[slot], jitter=self._slot_jitter[slot])
        self.step_done[k].record(cur)
        i = self.n_run % self.log.shape[0]
        self.d2h_stream.wait_event(self.step_done[k])   # ONE 48-byte device -> host copy per step, off the compute stream
        _lib.check(self._copy_async(self._log_ptr + i * self._log_row_bytes, self._set_ptr[k], self._log_row_bytes, 2, self._d2h_raw))
        self.d2h_done[k].record(self.d2h_stream)
        self.free[slot].record(cur)
        self.n_run += 1

    def losses(self) -> List[Dict[str, float]]:
        """Synchronise and decode the logged loss terms of the last min(n_run, log_capacity) steps."""
        torch.cuda.current_stream(self.tr.device).synchronize()
        self.d2h_stream.synchronize()
        c, tr = self.tr.conf, self.tr
        out = []
        cap = self.log.shape[0]
        for i in range(max(0, self.n_run - cap), self.n_run):
            r, t = self.log[i % cap, :8].view(torch.float32).tolist(), self.log[i % cap, 8:].tolist()
            S = max(t[0], 1)
            normal, mask, eik = r[1] / r[0], r[2] / (tr.n_patches * P), r[3] / (S * P)
            out.append(dict(loss=c["normal_weight"] * normal + c["mask_weight"] * mask + c["eikonal_weight"] * eik, normal=normal,
                            mask=mask, eikonal=eik, n_samples=t[0], overflow=t[2]))
        return out


def make_batch_struct(rays_o, rays_d, plane_n, near, far, v_inv, normal_gt, mask) -> SnbPatchBatch:
    """rays_o may be [N,3] or the reference's expanded [N,3,3,3] (only the centre origin is read)."""
    n = rays_d.shape[0]
    return SnbPatchBatch(n, *[0 if t is None else t.data_ptr() for t in (rays_o, rays_d, plane_n, near, far, v_inv, normal_gt, mask)])


class FusedTrainer:
    """Runner.train (exp_runner.py:147-210) on the fused kernels.  One process per GPU; with world_size > 1 (torch.distributed
    initialised) each rank draws its own patches (weak scaling) and the gradients of all ranks are summed, scaled by
    1 / world_size and applied identically everywhere: inside the step-tail kernel over NVLink peer memory (snb_train_tail_peer:
    peer loads, Adam state sharded by chunk owner, parameter broadcast), or by an NCCL all-reduce + replicated Adam with SNB_DP=nccl."""

    def __init__(self, dataset, conf: dict, device="cuda", seed: int = 0, samples_per_ray_cap: int = 320,
                 world_size: int = 1, rank: int = 0):
        self.ds, self.conf, self.device = dataset, conf, torch.device(device)
        self.world_size, self.rank = world_size, rank
        self.n_patches = int(conf["batch_size"])
        assert int(conf["patch_size"]) == 3, "3x3 patches (config/diligent.conf:29)"
        # 'dfd' (both shipped confs) and 'ad' (models/renderer.py:225-226, SDFNetwork.gradient with create_graph=True) both run fully fused:
        # 'ad' swaps the render stage for snb_render_fused_ad on analytic gradients (snb_sdf_grad_patch) and adds their double backward
        # (snb_sdf_grad_bwd_patch).  SNB_AD_FUSED=0 keeps the autograd route (forward_backward_ad: fused marcher / Adam around the drop-in
        # autograd operators) as the cross-check.
        self.gradient_method = conf.get("gradient_method", "dfd")
        if self.gradient_method not in ("dfd", "ad"):
            raise NotImplementedError(f"gradient_method {self.gradient_method!r}: 'dfd' and 'ad' are implemented ('fd' is a debugging aid of the reference)")
        self.ad_fused = self.gradient_method == "ad" and os.environ.get("SNB_AD_FUSED", "1") != "0"
        # the kernels hard-wire what both shipped confs select; anything else must fail here, not train silently on other maths
        if conf.get("loss_type", "l1") != "l2":     # exp_runner.py:61 defaults to 'l1' when the key is absent
            raise NotImplementedError(f"fused trainer implements loss_type 'l2' (config/diligent.conf, own_objects.conf); got {conf.get('loss_type', 'l1')!r}")
        if not float(conf.get("mask_weight", 0.0)) > 0.0:   # exp_runner.py:171-172: mask_weight == 0 switches to an all-ones mask
            raise NotImplementedError("fused trainer implements mask_weight > 0 (mask from the dataset); mask_weight == 0 uses an all-ones mask in the reference")
        # data parallel: gradient reduction + sharded Adam + parameter broadcast inside ONE kernel over NVLink peer memory
        # (SNB_DP=peer, default) or NCCL allreduce + replicated Adam (SNB_DP=nccl, also the fallback when symmetric memory is unavailable)
        self.peer_mode = False
        self.model = None
        if world_size > 1 and os.environ.get("SNB_DP", "peer") == "peer" and (self.gradient_method == "dfd" or self.ad_fused):
            import torch.distributed as dist
            ok = 1
            try:
                self.model = SDFModel(conf["encoding"], conf["sdf_network"]["bias"], conf["variance_init"], device=self.device,
                                      peer_group_factory=lambda nbytes: dp.PeerGroup(nbytes, self.device))
            except Exception as e:   # noqa: BLE001 -- any failure of the symmetric-memory setup selects the NCCL path, loudly
                import sys
                print(f"[supernormal_b200] rank {rank}: peer-memory data parallelism unavailable ({type(e).__name__}: {e}); using NCCL allreduce",
                      file=sys.stderr)
                ok, self.model = 0, None
            flag = torch.tensor([ok], device=self.device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)      # every rank takes the same path
            self.peer_mode = bool(flag.item())
            if not self.peer_mode:
                self.model = None
        if self.model is None:
            self.model = SDFModel(conf["encoding"], conf["sdf_network"]["bias"], conf["variance_init"], device=self.device)
        rm = conf["ray_marching"]
        self.start_step, self.end_step = float(rm["start_step_size"]), float(rm["end_step_size"])
        self.slop = (math.log10(self.start_step) - math.log10(self.end_step)) / conf["end_iter"]
        stride = int(2.0 / self.end_step) + 64
        self.buf = SampleBuffers(self.n_patches, samples_per_ray_cap, stride, self.model.n_levels, self.device)
        if self.ad_fused:    # analytic gradients of the sample starts and their seeds, f32 [9 * capacity, 3] each
            self.buf.grad = torch.zeros(P * self.buf.capacity * 3, device=self.device)
            self.buf.d_grad = torch.zeros(P * self.buf.capacity * 3, device=self.device)
        from .nerfacc_api import OccupancyGrid
        self.grid = OccupancyGrid([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], 128).to(self.device)
        self.seed = dp.rank_seed(seed, rank)   # every rank draws its own patches (weak scaling)
        self.fused_host = True        # one C-ABI call per phase instead of one per kernel
        self._ctx_cache: Dict[tuple, SnbTrainCtx] = {}
        self.lean = (self.peer_mode or os.environ.get("SNB_LEAN", "1") != "0") and (self.gradient_method == "dfd" or self.ad_fused)   # train_step: no prep_net / unfold_grads / sample_patches launches -- the step-tail kernel
                                      # (snb_train_tail) unfolds, runs Adam, folds the updated weights and pre-samples the next batch
        self.legacy_render = False    # per-kernel path only: render_fwd / patch_loss / render_bwd instead of render_fused
        self.device_sampler = True    # snb_sample_patches instead of the ATen-op gen_random_patches
        dv = self.device
        n, Pn = self.n_patches, P
        # two batch slots: the tail kernel of iteration `it` writes the batch of `it + 1` into the slot iteration `it` is not using
        self._own = [dict(rays_o=torch.zeros(n, 3, device=dv), rays_d=torch.zeros(n, Pn, 3, device=dv), plane_n=torch.zeros(n, 3, device=dv),
                          near=torch.zeros(n, device=dv), far=torch.zeros(n, device=dv), v_inv=torch.zeros(n, Pn, 9, device=dv),
                          normal_gt=torch.zeros(n, Pn, 3, device=dv), mask=torch.zeros(n, Pn, device=dv)) for _ in range(2)]
        self._own_jitter = [torch.zeros(n, device=dv) for _ in range(2)]
        self._slot = 0                 # slot of the batch sample_batch_device() last handed out
        self._presampled = (-1, 0)     # (iteration, slot) filled ahead by the tail kernel
        self.train_ids = torch.tensor(dataset.train_images, dtype=torch.int32, device=dv)
        # the C ABI reads dense fp32 tensors (torch.inverse may hand back column-major batches)
        # V_inverse (models/dataset_loader.py:114-137) is computed in closed form inside the sampler; SNB_VINV_TABLE=1 reads the dataset's
        # precomputed [n_images,H,W,3,3] table instead (36 B per pixel: 226 MB for DiLiGenT-MV) -- tests compare the two
        self.v_inverse_table = os.environ.get("SNB_VINV_TABLE", "0") == "1"
        self._ds_tensors = [t.to(dv, torch.float32).contiguous() for t in (dataset.normals, dataset.masks, dataset.intrinsics_all_inv, dataset.pose_all)]
        if self.v_inverse_table:
            self._ds_tensors.append(dataset.V_inverse_all.to(dv, torch.float32).contiguous())
        self.ds_struct = SnbDataset(dataset.n_images, dataset.H, dataset.W, len(dataset.train_images),
                                    *[t.data_ptr() for t in self._ds_tensors[:4]], self._ds_tensors[4].data_ptr() if self.v_inverse_table else None,
                                    self.train_ids.data_ptr())
        self._out_structs = [SnbBatchOut(*[ob[k].data_ptr() for k in ("rays_o", "rays_d", "plane_n", "near", "far", "v_inv", "normal_gt", "mask")],
                                         jit.data_ptr()) for ob, jit in zip(self._own, self._own_jitter)]
        self.occs_prev = torch.zeros(self.grid.num_cells, device=dv)
        self.occ_ws = torch.zeros(2, dtype=torch.float64, device=dv)
        self.iter_step = 0
        self.lr = float(conf["learning_rate"]) * self._lr_factor()   # Runner.train calls update_learning_rate() before the loop
                                                                      # (exp_runner.py:125): step 0 runs at lr = 0, only Adam's moments move
        self.gen = torch.Generator(device=self.device).manual_seed(seed + rank)
        self.np_rng = np.random.RandomState(seed + rank)
        self.last_batch = None
        if self.peer_mode:
            import torch.distributed as dist
            self._peer_struct = self.model.peer_struct()
            self._peer_epoch = 0
            torch.cuda.synchronize(self.device)
            dist.barrier()           # every rank's parameters are in place before anyone's first peer kernel

    @property
    def own_batch(self) -> Dict[str, torch.Tensor]:
        """The device-sampled batch of the current / last iteration."""
        return self._own[self._slot]

    @property
    def own_jitter(self) -> torch.Tensor:
        return self._own_jitter[self._slot]

    # -- schedule pieces -----------------------------------------------------------------------
    def step_size(self, it: int) -> float:
        return float(np.float32(10 ** (math.log10(self.start_step) - self.slop * it)))

    def _lr_factor(self) -> float:
        c = self.conf
        if self.iter_step < c["warm_up_end"]:
            return self.iter_step / c["warm_up_end"]
        prog = (self.iter_step - c["warm_up_end"]) / (c["end_iter"] - c["warm_up_end"])
        return (math.cos(math.pi * prog) + 1.0) * 0.5 * (1 - c["learning_rate_alpha"]) + c["learning_rate_alpha"]

    def update_occupancy(self, it: int):
        rm = self.conf["ray_marching"]
        if not self.lean or self.model.net_stale:   # the lean step keeps `net` folded (snb_train_tail)
            self.model.prep()
        if self.fused_host:
            net = self.model.net_struct()
            binary = self.grid._binary
            call("snb_occgrid_update_fused", C.byref(net), *self.grid._res, ptr(self.grid.roi_aabb), int(it < 256), 0.95,
                 float(rm["occ_threshold"]), self.seed, it, ptr(self.grid.occs), ptr(self.occs_prev), ptr(binary.view(torch.uint8)),
                 ptr(self.occ_ws))
        else:
            self.grid._update(step=it, occ_eval_fn=lambda x: self.model.sdf(x, mode=1), occ_thre=rm["occ_threshold"])

    def sample_batch_device(self, it: int):
        """Device-side sampler: self.own_batch / self.own_jitter for iteration `it` -- the slot the previous iteration's tail
        kernel already filled, or one launch now.  Counter-based RNG keyed by (seed, it): the batch is the same either way."""
        if self._presampled[0] == it:
            self._slot = self._presampled[1]
        else:
            call("snb_sample_patches", C.byref(self.ds_struct), self.n_patches, self.seed, it, C.byref(self._out_structs[self._slot]))
        self._presampled = (-1, 0)
        return self.own_batch, self.own_jitter

    def _ctx(self, batch: dict, jitter) -> SnbTrainCtx:
        """The fused host calls' argument block.  Rebuilding it is ~35 data_ptr() calls + two ctypes structs, twice per step; every pointer
        in it is a fixed buffer, so it is cached on what can change between steps: the batch buffers (device-sampler slots / feeder
        slots), the jitter, the accumulator set (SampleBuffers.flip) and the live level count."""
        m = self.model
        key = (batch["rays_o"].data_ptr(), batch["mask"].data_ptr(), 0 if jitter is None else jitter.data_ptr(), self.buf.set_idx,
               m.n_active, self.grid.binary.data_ptr(), m.flat.data_ptr(), m.grad.data_ptr(), m.table_f16.data_ptr())   # + the buffers a caller may swap
        hit = self._ctx_cache.get(key)
        if hit is not None:
            return hit
        if len(self._ctx_cache) > 64:
            self._ctx_cache.clear()
        ctx = self._ctx_cache[key] = self._build_ctx(batch, jitter)
        return ctx

    def _build_ctx(self, batch: dict, jitter) -> SnbTrainCtx:
        m, b = self.model, self.buf
        bs = make_batch_struct(batch["rays_o"], batch["rays_d"], batch["plane_n"], batch["near"], batch["far"],
                               batch["v_inv"], batch["normal_gt"], batch["mask"])
        g = self.grid
        return SnbTrainCtx(bs, b.struct, m.net_struct(), m.n_levels, SMALL_PAD, m.flat.data_ptr(), m.grad.data_ptr(),
                           m.exp_avg.data_ptr(), m.exp_avg_sq.data_ptr(), m.net_grad.data_ptr(), b.sdf.data_ptr(), b.feats.data_ptr(),
                           b.d_sdf0.data_ptr(), b.d_sdf1.data_ptr(), b.comp.data_ptr(), b.wsum.data_ptr(), b.dcomp.data_ptr(),
                           b.dwsum.data_ptr(), b.stats.data_ptr(), 0 if jitter is None else jitter.data_ptr(), g.roi_aabb.data_ptr(),
                           g.binary.data_ptr(), *g._res, b.bwd_ws.data_ptr(), b.bwd_ws_bytes)

    def sample_batch(self):
        c = self.conf
        o, d, pn, vinv, nrm, msk = self.ds.gen_random_patches(self.n_patches, c["patch_size"], c["patch_size"],
                                                              generator=self.gen, np_rng=self.np_rng)
        near, far = self.ds.near_far_from_sphere(o[:, 1, 1], d[:, 1, 1])
        return dict(rays_o=o[:, 1, 1].contiguous(), rays_d=d.view(-1, P, 3), plane_n=pn, near=near.contiguous(),
                    far=far.contiguous(), v_inv=vinv.view(-1, P, 9), normal_gt=nrm.view(-1, P, 3), mask=msk.view(-1, P).contiguous())

    # -- one iteration ---------------------------------------------------------------------------
    def forward_backward(self, batch: dict, step_size: float, jitter: Optional[torch.Tensor], lean: bool = False):
        """march -> sdf -> render -> loss -> backward; leaves gradients in model.grad (small + table).
        lean=True (what train_step uses): `net` is already folded and net_grad zero (left by the previous tail kernel), the
        loss accumulators are reset by the compaction kernel, and the MLP gradient stays in model.net_grad w.r.t. the FOLDED
        weights -- optimizer_step's tail kernel unfolds it."""
        m, b = self.model, self.buf
        c = self.conf
        assert lean or not self.peer_mode, "peer-memory data parallelism runs the lean step only (net_grad aliases the gradient buffer)"
        self.last_batch = batch
        self._grads_folded = lean
        if lean and m.net_stale:
            m.prep()
            m.net_grad.zero_()
            m.net_stale = False
        if self.fused_host and self.ad_fused and lean:
            ctx = self._ctx(batch, jitter)
            call("snb_train_fwd_bwd_lean_ad", C.byref(ctx), float(step_size), 1e-8, float(c["normal_weight"]), float(c["mask_weight"]),
                 float(c["eikonal_weight"]), ptr(b.grad), ptr(b.d_grad))
            _lib.LAUNCH_COUNT += 7 - _lib.KERNELS.get("snb_train_fwd_bwd_lean_ad", 1) + (1 if m.n_active > 4 else 0)
            return
        if self.fused_host and not self.ad_fused:
            ctx = self._ctx(batch, jitter)
            call("snb_train_fwd_bwd_lean" if lean else "snb_train_fwd_bwd", C.byref(ctx), float(step_size), 1e-8, float(c["normal_weight"]),
                 float(c["mask_weight"]), float(c["eikonal_weight"]))
            if m.n_active > 4 and os.environ.get("SNB_BWD_SPLIT", "1") != "0" and os.environ.get("SNB_BWD_UMMA", "1") != "0":
                _lib.LAUNCH_COUNT += 1   # beyond 4 live levels the backward is two launches (MLP backward + table scatter)
            return
        bs = make_batch_struct(batch["rays_o"], batch["rays_d"], batch["plane_n"], batch["near"], batch["far"],
                               batch["v_inv"], batch["normal_gt"], batch["mask"])
        if not lean:
            m.prep(batch["mask"], b.stats)
        net = m.net_struct()
        rb, rn, rs = C.byref(bs), C.byref(net), C.byref(b.struct)
        res = self.grid._res
        grid_u8 = self.grid.binary.view(torch.uint8)
        call("snb_march_visible", rb, rn, ptr(self.grid.roi_aabb), *res, ptr(grid_u8), float(step_size), ptr(jitter), 1e-8, rs)
        if lean:
            call("snb_compact_samples_stats", self.n_patches, rs, batch["mask"].numel(), ptr(batch["mask"]), ptr(b.stats))
        else:
            call("snb_compact_samples", self.n_patches, rs)
        call("snb_sdf_fwd_patch", rb, rn, rs, ptr(b.sdf), ptr(b.feats))
        if self.legacy_render:   # the three serial-chain kernels the single-launch render stage replaced (kept for cross-checks)
            call("snb_render_fwd", rb, rn, rs, ptr(b.sdf), ptr(b.comp), ptr(b.wsum), None, None, ptr(b.stats))
            call("snb_patch_loss", rb, ptr(b.comp), ptr(b.wsum), float(c["normal_weight"]), float(c["mask_weight"]), ptr(b.stats),
                 ptr(b.dcomp), ptr(b.dwsum))
            call("snb_render_bwd", rb, rn, rs, ptr(b.sdf), ptr(b.comp), ptr(b.wsum), ptr(b.dcomp), ptr(b.dwsum), None,
                 float(c["eikonal_weight"]), ptr(b.d_sdf0), ptr(b.d_sdf1), ptr(b.stats))
        elif self.ad_fused:
            call("snb_sdf_grad_patch", rb, rn, rs, ptr(b.grad))
            call("snb_render_fused_ad", rb, rn, rs, ptr(b.sdf), ptr(b.grad), float(c["normal_weight"]), float(c["mask_weight"]),
                 float(c["eikonal_weight"]), ptr(b.comp), ptr(b.wsum), ptr(b.d_sdf0), ptr(b.d_sdf1), ptr(b.d_grad), ptr(b.stats))
        else:
            call("snb_render_fused", rb, rn, rs, ptr(b.sdf), float(c["normal_weight"]), float(c["mask_weight"]), float(c["eikonal_weight"]),
                 ptr(b.comp), ptr(b.wsum), ptr(b.d_sdf0), ptr(b.d_sdf1), ptr(b.stats))
        if not lean:
            m.net_grad.zero_()
        call("snb_sdf_bwd_patch_ws", rb, rn, rs, ptr(b.feats), ptr(b.d_sdf0), ptr(b.d_sdf1), ptr(m.grad[SMALL_PAD:]), ptr(m.net_grad),
             ptr(b.bwd_ws), b.bwd_ws_bytes)
        if self.ad_fused:
            call("snb_sdf_grad_bwd_patch", rb, rn, rs, ptr(b.feats), ptr(b.d_grad), ptr(m.grad[SMALL_PAD:]), ptr(m.net_grad))
        if not lean:
            call("snb_unfold_grads", m.n_levels, ptr(m.small), ptr(m.net_grad), ptr(b.stats), ptr(m.grad))

    def forward_backward_ad(self, batch: dict, step_size: float, jitter: Optional[torch.Tensor]):
        """One iteration's forward + backward with ANALYTIC normals (gradient_method = 'ad'; models/renderer.py:225-226 and
        SDFNetwork.gradient, models/fields.py:107-119).  Fused: ray marching + visibility cut + sample compaction (snb_march_visible,
        snb_compact_samples) and, afterwards, Adam.  Through the drop-in operators under torch autograd: the hash-grid encoding
        (tcnn_api: first derivative w.r.t. x and the double backward into table and dL/dy, snb_hashgrid_bwd_input /
        snb_hashgrid_bwd_bwd_input), the fp32 MLP, NeuS alpha (models/renderer.py:164-179), patch weights and accumulation
        (nerfacc_api) and the losses (exp_runner.py:191-203).  Leaves the gradients w.r.t. (v, g, b, variance | table) in model.grad
        and the loss terms in buf.stats, like forward_backward(lean=False); one host read-back (the sample counts)."""
        import torch.nn.functional as F
        from . import nerfacc_api as na
        from . import tcnn_api
        m, b, c = self.model, self.buf, self.conf
        self.last_batch = batch
        self._grads_folded = False
        n = self.n_patches
        m.prep(batch["mask"], b.stats)      # the marcher's no-grad SDF uses the folded weights
        m.net_stale = False
        bs = make_batch_struct(batch["rays_o"], batch["rays_d"], batch["plane_n"], batch["near"], batch["far"], batch["v_inv"],
                               batch["normal_gt"], batch["mask"])
        net = m.net_struct()
        rs = C.byref(b.struct)
        call("snb_march_visible", C.byref(bs), C.byref(net), ptr(self.grid.roi_aabb), *self.grid._res, ptr(self.grid.binary.view(torch.uint8)),
             float(step_size), ptr(jitter), 1e-8, rs)
        call("snb_compact_samples", n, rs)
        S = int(b.totals[0].item())
        m.grad.zero_()
        if S == 0:      # models/renderer.py:137-140: nothing to render, the reference skips the iteration
            b.stats[1:].zero_()
            return
        pidx = b.patch_idx[:S].long()
        t0c, t1c = b.t0[:S, None], b.t1[:S, None]
        o = batch["rays_o"]                                         # [n,3] (the patch's camera centre)
        d = batch["rays_d"].view(n, P, 3)
        pn = batch["plane_n"]
        # ---- differentiable parameters: leaves over the flat buffer
        so = m._small_offsets()
        small = m.small.detach().clone().requires_grad_(True)
        table = m.table.detach().requires_grad_(True)
        v0, g0, b0 = small[so["v0"]:so["g0"]].view(H, m.d_in), small[so["g0"]:so["b0"]].view(H, 1), small[so["b0"]:so["v1"]]
        v1, g1, b1, var = small[so["v1"]:so["g1"]].view(1, H), small[so["g1"]:so["g1"] + 1].view(1, 1), small[so["b1"]:so["b1"] + 1], small[so["var"]]
        W0, W1 = g0 * v0 / v0.norm(dim=1, keepdim=True), g1 * v1 / v1.norm(dim=1, keepdim=True)     # nn.utils.weight_norm, models/fields.py:66-67
        if getattr(self, "_ad_enc", None) is None:
            self._ad_enc = tcnn_api.Encoding(3, dict(c["encoding"])).to(self.device)
            self._ad_enc.params.requires_grad_(False)
        enc = self._ad_enc
        enc.n_active_levels = m.n_active                             # levels >= bindwidth contribute exact zeros (models/fields.py:81-83)
        # the gathers read the fp16 copy the optimizer kernel maintains (tcnn_api would otherwise recast -- and its cache cannot see the
        # in-place updates of snb_train_optim)
        enc._cache_table, enc._cache_key = m.table_f16, (table.data_ptr(), table._version, table.device)

        def sdf_fn(x):                                               # SDFNetwork.forward, models/fields.py:76-99
            feat = tcnn_api._Encode.apply(enc, x, table).to(torch.float32)
            h = F.softplus(F.linear(torch.cat([x, feat], 1), W0, b0), beta=100)
            return F.linear(h, W1, b1)

        with torch.no_grad():                                        # plane fan-out, models/renderer.py:146-159
            num = (d[:, P // 2] * pn).sum(-1, keepdim=True)[pidx][:, None, :]
            den = (d * pn[:, None, :]).sum(-1, keepdim=True)[pidx]
            t0, t1 = t0c[:, None, :] * num / den, t1c[:, None, :] * num / den
            nxt = torch.cat([t0c[1:], t0c[-1:]], 0)
            dm = ((t1c - nxt) != 0)[:, 0]
            p0 = o[pidx][:, None, :] + d[pidx] * t0                  # [S,9,3]
            p1 = o[pidx][:, None, :] + d[pidx] * t1
            pos_all = torch.cat([p0, p1[dm]], 0)
        sdf_all = sdf_fn(pos_all.reshape(-1, 3)).reshape(-1, P, 1)
        s0 = sdf_all[:S]
        s1 = torch.cat([s0[1:], s0[-1:]], 0).clone()                 # models/renderer.py:164-169
        s1[dm] = sdf_all[S:]
        inv_s = torch.exp(var * 10.0).clip(1e-6, 1e6)
        cdf0, cdf1 = torch.sigmoid(s0 * inv_s), torch.sigmoid(s1 * inv_s)
        alpha = ((cdf0 - cdf1 + 1e-5) / (cdf0 + 1e-5)).clip(0.0, 1.0)
        w = na.render_weight_from_alpha_patch_based(alpha, pidx)
        with torch.enable_grad():                                    # SDFNetwork.gradient, models/fields.py:107-119
            x = p0.reshape(-1, 3).detach().requires_grad_(True)
            y = sdf_fn(x)
            grads = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True, only_inputs=True)[0].reshape(S, P, 3)
        wsum = na.accumulate_along_rays_patch_based(w, pidx, n_patches=n)                       # [n,9,1]
        comp = na.accumulate_along_rays_patch_based(w, pidx, values=grads, n_patches=n)         # [n,9,3]
        # ---- losses, exp_runner.py:169-203 (l2)
        mask = (batch["mask"].view(n, P, 1) > 0.5).float()
        mask_sum = mask.sum() + 1e-5
        err = (comp - batch["normal_gt"].view(n, P, 3)) * mask
        normal_loss = (err ** 2).sum() / mask_sum
        eik = ((torch.linalg.norm(grads, ord=2, dim=-1) - 1.0) ** 2).mean()
        mloss = F.binary_cross_entropy(wsum.clip(1e-5, 1.0 - 1e-5), mask)
        loss = float(c["normal_weight"]) * normal_loss + float(c["mask_weight"]) * mloss + float(c["eikonal_weight"]) * eik
        loss.backward()
        with torch.no_grad():
            m.grad[:m.n_small] = small.grad
            m.grad[SMALL_PAD:] = table.grad
            b.comp.view(n, P, 3).copy_(comp)
            b.wsum.view(n, P).copy_(wsum[..., 0])
            # the layout loss_terms() decodes: [mask_sum | normal * mask_sum | bce * n * 9 | eikonal * S * 9 | d loss / d inv_s ...]
            b.stats[0] = mask_sum
            b.stats[1] = normal_loss.detach() * mask_sum
            b.stats[2] = mloss.detach() * (n * P)
            b.stats[3] = eik.detach() * (S * P)

    def optimizer_step(self, presample_next: bool = False):
        """Adam (exp_runner.py:205-207).  After a lean forward_backward: ONE launch (snb_train_tail) that also unfolds the MLP
        gradient, folds the updated weights for the next forward and, with presample_next, draws the next iteration's batch."""
        m = self.model
        n_live = dp.live_numel(SMALL_PAD, m.offsets, m.n_active)  # zero-gradient levels are exact no-ops for Adam without weight decay
        t = self.iter_step + 1
        ctx = self._ctx(self.last_batch, None)
        if not getattr(self, "_grads_folded", False):
            gscale = dp.allreduce_live_gradients(m.grad, n_live, self.world_size)
            call("snb_train_optim", C.byref(ctx), float(self.lr), t, gscale)   # one Adam sweep: MLP block + live table levels + fp16 refresh
            m.net_stale = True
            return
        nxt = 1 - self._slot
        if self.peer_mode:   # reduction over NVLink peer memory, sharded Adam, parameter broadcast, fold, next batch: one launch
            self._peer_epoch += 1                         # barrier epoch: a launch counter, NOT iter_step (load_state_dict rewinds that)
            self._peer_struct.epoch = self._peer_epoch
            if self._peer_epoch % 256 == 0:
                self.check_peer_error()                   # a timed-out cross-GPU wait must stop the run, not train on unreduced gradients
            if presample_next:
                call("snb_train_tail_peer", C.byref(ctx), C.byref(self._peer_struct), float(self.lr), t, C.byref(self.ds_struct), self.n_patches,
                     self.seed, self.iter_step + 1, C.byref(self._out_structs[nxt]))
                self._presampled = (self.iter_step + 1, nxt)
            else:
                call("snb_train_tail_peer", C.byref(ctx), C.byref(self._peer_struct), float(self.lr), t, None, 0, 0, 0, None)
            self._grads_folded = False
            return
        unfolded = 0
        gscale = 1.0
        if self.world_size > 1:   # the allreduce works on (v, g, b, variance) gradients: unfold first, tail skips its unfold
            call("snb_unfold_grads", m.n_levels, ptr(m.small), ptr(m.net_grad), ptr(self.buf.stats), ptr(m.grad))
            gscale = dp.allreduce_live_gradients(m.grad, n_live, self.world_size)
            unfolded = 1
        if presample_next:
            call("snb_train_tail", C.byref(ctx), float(self.lr), t, gscale, unfolded, C.byref(self.ds_struct), self.n_patches, self.seed,
                 self.iter_step + 1, C.byref(self._out_structs[nxt]))
            self._presampled = (self.iter_step + 1, nxt)
        else:
            call("snb_train_tail", C.byref(ctx), float(self.lr), t, gscale, unfolded, None, 0, 0, 0, None)
        self._grads_folded = False

    def train_step(self, batch: Optional[dict] = None, jitter: Optional[torch.Tensor] = None):
        c, rm = self.conf, self.conf["ray_marching"]
        it = self.iter_step
        if it % rm["occ_update_freq"] == 0:
            self.update_occupancy(it)
        if it % c["increase_bindwidth_every"] == 0:
            self.model.n_active = min(self.model.n_active + 1, self.model.n_levels)
        own = batch is None and self.device_sampler
        if own:
            batch, jitter = self.sample_batch_device(it)
        if batch is None:
            batch = self.sample_batch()
        if jitter is None:
            jitter = torch.rand(self.n_patches, device=self.device, generator=self.gen)
        if self.gradient_method == "ad" and not self.ad_fused:
            self.forward_backward_ad(batch, self.step_size(it), jitter)
        else:
            self.forward_backward(batch, self.step_size(it), jitter, lean=self.lean)
        self.optimizer_step(presample_next=own and self.lean)
        self.iter_step += 1
        self.lr = self.conf["learning_rate"] * self._lr_factor()

    # -- checkpoint / resume -------------------------------------------------------
    def _param_slices(self):
        """(name, offset into the flat buffers, shape) in the order of `list(sdf_network.parameters()) + list(deviation_network.parameters())`
        (exp_runner.py:86-91): tcnn's `params`, then per Linear `bias, weight_g, weight_v` (nn.utils.weight_norm re-registers g and v behind
        the bias, models/fields.py:66-67), then `variance`.  These are the indices torch.optim.Adam.state_dict() numbers its state by."""
        m = self.model
        o = m._small_offsets()
        return [("encoding.params", SMALL_PAD, (m.n_table,)),
                ("lin0.bias", o["b0"], (H,)), ("lin0.weight_g", o["g0"], (H, 1)), ("lin0.weight_v", o["v0"], (H, m.d_in)),
                ("lin1.bias", o["b1"], (1,)), ("lin1.weight_g", o["g1"], (1, 1)), ("lin1.weight_v", o["v1"], (1, H)),
                ("variance", o["var"], ())]

    def state_dict(self) -> dict:
        """What Runner.save_checkpoint stores (exp_runner.py:306-315): `sdf_network_fine`, `variance_network_fine`, `optimizer` =
        torch.optim.Adam.state_dict() layout (`state[i] = {step, exp_avg, exp_avg_sq}` per parameter tensor in the reference's
        parameter order, `param_groups`), `iter_step` -- the reference's own load_checkpoint (exp_runner.py:298-304) reads it.  Two
        extra keys the reference forgets and a resume needs are added (the reference ignores unknown keys): `occupancy_grid` (the
        reference restarts from an all-empty grid) and `n_active`.  Collective in a data-parallel run with the peer-memory tail:
        the table's Adam state is sharded by chunk owner and is summed back together (non-owned chunks are exactly zero)."""
        m = self.model
        ea, eas = m.exp_avg.clone(), m.exp_avg_sq.clone()
        if self.peer_mode:
            import torch.distributed as dist
            for t in (ea, eas):
                small = t[:SMALL_PAD].clone()    # the MLP block is updated identically on every rank: not summed
                t[:SMALL_PAD] = 0
                dist.all_reduce(t)
                t[:SMALL_PAD] = small
        sd = m.reference_state_dict()
        state = {}
        if self.iter_step > 0:                   # torch's Adam creates its state lazily at the first step()
            for i, (_, off, shape) in enumerate(self._param_slices()):
                n = int(np.prod(shape)) if len(shape) else 1
                state[i] = {"step": torch.tensor(float(self.iter_step)), "exp_avg": ea[off:off + n].clone().view(shape),
                            "exp_avg_sq": eas[off:off + n].clone().view(shape)}
        group = dict(torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=float(self.conf["learning_rate"])).state_dict()["param_groups"][0])
        group["lr"] = float(self.lr)             # exp_runner.py:278-279 writes the scheduled value into the group
        group["params"] = list(range(len(self._param_slices())))
        sd["optimizer"] = {"state": state, "param_groups": [group]}
        sd["iter_step"] = self.iter_step
        sd["n_active"] = m.n_active
        sd["occupancy_grid"] = {"occs": self.grid.occs.clone(), "binary": self.grid._binary.clone()}
        return sd

    @torch.no_grad()
    def rebuild_occupancy(self) -> None:
        """Occupancy grid from the current SDF alone (a reference checkpoint does not carry one): one all-cells sweep from an empty
        grid, the form every_n_step uses while step < 256 (NA/grid.py:207-239)."""
        self.grid.occs.zero_()
        self.occs_prev.zero_()
        self.grid._binary.zero_()
        self.occ_ws.zero_()
        self.update_occupancy(0)

    @torch.no_grad()
    def load_state_dict(self, sd: dict) -> None:
        """Resume from state_dict() or from a checkpoint written by the reference's save_checkpoint: parameters, Adam moments,
        schedule position (step size, learning rate and the number of active levels are functions of iter_step) and the occupancy
        grid (rebuilt from the SDF when the checkpoint has none).  The patch stream is counter-based (seed, iteration), so a resumed
        run draws the batches the uninterrupted run would have drawn."""
        m = self.model
        self._ctx_cache.clear()
        m.load_reference_state_dict(sd)
        m.exp_avg.zero_()
        m.exp_avg_sq.zero_()
        state = sd["optimizer"]["state"]
        for i, (_, off, shape) in enumerate(self._param_slices()):
            st = state.get(i, state.get(str(i)))
            if st is None:
                continue
            n = int(np.prod(shape)) if len(shape) else 1
            m.exp_avg[off:off + n] = st["exp_avg"].to(self.device, torch.float32).flatten()
            m.exp_avg_sq[off:off + n] = st["exp_avg_sq"].to(self.device, torch.float32).flatten()
        if self.peer_mode:   # keep only the chunks this rank owns
            mine = dp.owner_mask(m.n_table, self.rank, self.world_size, self.device)
            m.exp_avg[SMALL_PAD:] *= mine
            m.exp_avg_sq[SMALL_PAD:] *= mine
        m.grad.zero_()
        self.iter_step = int(sd["iter_step"])
        every = int(self.conf["increase_bindwidth_every"])
        # exp_runner.py:158-159: one more level at every iteration with iter_step % every == 0, starting with iteration 0
        m.n_active = int(sd["n_active"]) if "n_active" in sd else min((self.iter_step + every - 1) // every, m.n_levels)
        self.lr = float(self.conf["learning_rate"]) * self._lr_factor()
        self._presampled = (-1, 0)
        if "occupancy_grid" in sd:
            self.grid.occs.copy_(sd["occupancy_grid"]["occs"].to(self.device))
            self.grid._binary.copy_(sd["occupancy_grid"]["binary"].to(self.device))
            self.occ_ws.zero_()     # workspace of the fused occupancy update: [sum of occs (f64) | number of occupied cells (u64)]
            self.occ_ws.view(torch.int64)[1] = int(self.grid._binary.sum().item())
        else:
            self.rebuild_occupancy()
        if self.peer_mode:
            import torch.distributed as dist
            torch.cuda.synchronize(self.device)
            dist.barrier()

    # -- host-fed training (the end-to-end path: batches arrive in HOST memory, losses go back to the host) -----------
    def host_feeder(self, depth: int = 2, log_capacity: int = 4096) -> "HostBatchFeeder":
        return HostBatchFeeder(self, depth, log_capacity)

    # -- host-side readbacks (sync) ---------------------------------------------------------------
    def check_peer_error(self) -> None:
        """Raises if a cross-GPU wait inside the peer tail kernel timed out (a rank never arrived)."""
        if self.peer_mode:
            code = int(self.model.peer_flags[16].item())
            if code:
                raise RuntimeError(f"snb_train_tail_peer: wait timed out (code {code}: 1xx/2xx start barrier, 300 local blocks, 4xx done barrier)")

    def loss_terms(self) -> Dict[str, float]:
        self.check_peer_error()
        st = self.buf.stats.cpu().tolist()
        tot = self.buf.totals.cpu().tolist()
        if tot[2]:   # set by the compaction kernel when a sample list was clipped to its capacity (or by a failed tcgen05 wait)
            import warnings
            warnings.warn(f"FusedTrainer: sample capacity exceeded at iteration {self.iter_step - 1} (samples_per_ray_cap = "
                          f"{self.buf.capacity // self.n_patches}); the step ran on truncated sample lists -- raise samples_per_ray_cap", RuntimeWarning)
        S = max(tot[0], 1)
        c = self.conf
        normal = st[1] / st[0]
        mask = st[2] / (self.n_patches * P)
        eik = st[3] / (S * P)
        return dict(loss=c["normal_weight"] * normal + c["mask_weight"] * mask + c["eikonal_weight"] * eik, normal=normal,
                    mask=mask, eikonal=eik, n_samples=tot[0], n_ends=tot[1], overflow=tot[2], samples_per_ray=tot[0] / self.n_patches)

    # -- measurement helpers ----------------------------------------------------------------------
    def algorithmic_bytes(self, name: str, S: int, E: int) -> Optional[int]:
        """HBM bytes one launch must move (SURVEY.md §8d per-unit figures x units), None if not modelled."""
        m = self.model
        M = P * (S + E)
        na = m.n_active
        if name == "snb_sdf_fwd_patch":      # write sdf (4 B) + kept features (4 B/level); read packed samples (12 B each)
            return M * (4 + 4 * na) + S * 12
        if name in ("snb_sdf_bwd_patch", "snb_sdf_bwd_patch_ws"):      # read kept features (4 B/level) + positions' inputs + seeds (d_sdf0, d_sdf1); the table-gradient
            return M * (4 * na + 4) + 2 * P * S * 4   # reductions are L2 traffic (l2_bytes), SURVEY.md §8d
        if name in ("snb_train_optim", "snb_train_tail"):   # p, g, m, v read + p, m, v, g(zero) write + fp16 copy
            return (SMALL_PAD + 2 * m.offsets[na]) * (16 + 16) + 2 * m.offsets[na] * 2
        if name == "snb_march_visible":      # 40 B in per ray + 8 B per emitted sample (scratch)
            return self.n_patches * 40 + S * 8
        if name == "snb_render_fused":       # sdf read + d_sdf0/d_sdf1 written per point, packed samples, per-ray constants and outputs
            return P * S * 16 + P * E * 4 + S * 12 + self.n_patches * P * (36 + 12 + 12 + 4 + 16)
        return None

    def l2_bytes(self, name: str, S: int, E: int) -> Optional[int]:
        """Useful L2 gather / reduction bytes of one launch (L * 8 corners * 4 B per point gathered, 8 B per corner reduced)."""
        M, na = P * (S + E), self.model.n_active
        if name == "snb_sdf_fwd_patch":
            return M * na * 8 * 4
        if name in ("snb_sdf_bwd_patch", "snb_sdf_bwd_patch_ws"):
            return M * na * 8 * 8
        return None

    def profile_kernels(self, steps: int = 20) -> dict:
        """Per-entry-point device time (CUDA events on the launching stream) over `steps` real iterations."""
        _lib.PROFILE = []
        acc = torch.zeros(2, dtype=torch.int64, device=self.device)   # sample / end counts summed on the device: no host sync per
        saved, self.fused_host = self.fused_host, False               # step (data-parallel ranks would drift apart); per-kernel entry
        try:                                                           # points so each launch can be bracketed
            for _ in range(steps):
                self.train_step()
                acc += self.buf.totals[:2]
            torch.cuda.synchronize()
            S_acc, E_acc = (int(v) for v in acc.tolist())
            rec = _lib.PROFILE
        finally:
            _lib.PROFILE = None
            self.fused_host = saved
        agg: Dict[str, list] = {}
        for name, e0, e1 in rec:
            agg.setdefault(name, []).append(e0.elapsed_time(e1) * 1e3)
        per_step = {k: sum(v) / steps for k, v in agg.items()}
        total = sum(per_step.values())
        out = {"us_per_step": {k: round(v, 2) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])},
               "sum_us_per_step": round(total, 2), "avg_samples": S_acc / steps, "avg_ends": E_acc / steps,
               "n_active": self.model.n_active}
        if self.peer_mode:
            out["note"] = "snb_train_tail_peer includes the in-kernel wait for the slowest rank (the two cross-GPU barriers)"
        S, E = int(S_acc / steps), int(E_acc / steps)
        out["hbm_model"] = {}
        for name in sorted(per_step, key=lambda k: -per_step[k]):
            nb = self.algorithmic_bytes(name, S, E)
            if nb:
                us = sum(agg[name]) / len(agg[name])
                calls = len(agg[name]) / steps
                rec = {"name": name, "us": us, "bytes": nb, "gbs": nb / (us * 1e-6) / 1e9, "share": per_step[name] / total,
                       "launches_per_step": calls}
                out["hbm_model"][name] = {"us": round(us, 2), "algorithmic_bytes": nb, "gbs": round(rec["gbs"], 1)}
                l2 = self.l2_bytes(name, S, E)
                if l2:
                    out["hbm_model"][name].update(l2_useful_bytes=l2, l2_useful_gbs=round(l2 / (us * 1e-6) / 1e9, 1))
                if "dominant" not in out:
                    out["dominant"] = rec
        return out
