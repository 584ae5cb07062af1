#!/usr/bin/env python
"""bench.py -- patch-rays/sec per training step (fwd+bwd+Adam) of the patch-based NeuS hot path.

Workload (BASELINE.json configs[1]): diligent.conf-shaped training -- synthetic analytic-sphere normal
maps, 20 views 612x512, 2048 patches of 3x3 rays per step, 14-level hash grid (T=2^19), the full
schedule (step size 1e-2 -> 1e-3, level activation every 350 it, occupancy update every 8 it, LR warm-up
+ cosine).  A "step" is one iteration of that schedule starting from random init: W warm-up iterations,
then exactly K timed ones.  `value` samples patches on the device; `e2e` feeds every step's patch batch
from pinned HOST memory and reads the loss terms back every step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "patch-rays/sec per train step (fwd+bwd)"
UNIT = "patch-rays/s"
WORKLOAD = "diligent.conf-shaped training: 20 views 612x512 synthetic sphere normals, 2048 patches x 3x3 rays/step, 14-level hash grid T=2^19, full schedule from random init"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).  The poller is started ahead of
    the warm-up (nvidia-smi takes a few hundred ms to emit its first row); rows are time-stamped on arrival and only those
    that fall inside [mark_start(), mark_end()] are summarised."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None
        self.nv_rows, self._stop, self.nv_thread = [], threading.Event(), None

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if ids and all(v.isdigit() for v in ids) and self.gpu < len(ids):
            return int(ids[self.gpu])
        return self.gpu

    def _nvml_loop(self):
        """In-process NVML polling (every 10 ms): the same fields as the nvidia-smi poller without its process start-up and
        per-row latency, which at 8 ranks can exceed a 300 ms timed region.  Any NVML failure leaves the nvidia-smi rows."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
            while not self._stop.is_set():
                sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = int(get_reasons(h))
                self.nv_rows.append((time.time(), sm, mx, [k for k, b in bits.items() if r & b]))
                time.sleep(0.01)
        except Exception:
            pass

    def start(self):
        self.nv_thread = threading.Thread(target=self._nvml_loop, daemon=True)
        self.nv_thread.start()
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        self._stop.set()
        if self.nv_thread is not None:
            self.nv_thread.join(timeout=1)
        if self.proc:
            time.sleep(0.12)          # let the row that covers the end of the region arrive
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        nv = [r for r in self.nv_rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or r[0])]
        if nv:
            return {"sm_mhz": float(np.median([r[1] for r in nv])), "sm_max_mhz": max(r[2] for r in nv),
                    "reasons": sorted({x for r in nv for x in r[3]}), "samples": len(nv), "window": "timed region", "source": "nvml"}
        ok = [r for ts, r in self.rows if len(r) >= 9 and self.t0 is not None and self.t0 <= ts <= (self.t1 or ts) + 0.06]
        where = "timed region"
        if not ok:   # region shorter than the polling period: take the rows closest to it
            ok = [r for ts, r in self.rows if len(r) >= 9 and self.t0 is not None and abs(ts - self.t0) < 0.5]
            where = "within 0.5 s of the timed region"
        sm = [float(r[1]) for r in ok if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in ok if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in ok:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": where}


def cpu_port_run(steps, warmup, n_patches=64, threads=None, budget_s=None):
    """The oracle (PyTorch-CPU port of the reference operators) on a bounded sample of the workload."""
    from oracle import torch_ops as T
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    ds = SyntheticDataset(SyntheticScene(), device="cpu")
    conf = dict(DILIGENT_CONF, batch_size=n_patches)
    tr = T.Trainer(ds, conf, seed=0)
    # bounded sample: the 2M-cell occupancy sweep of iteration 0 is replaced by its analytic result for the
    # geometric-init sphere (occupied where sdf < ~0.03, i.e. r < 0.63); timed steps skip grid updates.
    r = torch.arange(128).float().add(0.5).div(64).sub(1)
    gx, gy, gz = torch.meshgrid(r, r, r, indexing="ij")
    tr.renderer.occupancy_grid.binary = (gx ** 2 + gy ** 2 + gz ** 2).sqrt() < 0.63
    tr.renderer.occupancy_grid.every_n_step = lambda *a, **k: None
    tr.sdf.bindwidth = 0
    for _ in range(warmup):
        tr.step()
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        tr.step()
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:   # bounded sample: stop after the step that crosses the budget
            break
    steps = done
    dt = time.perf_counter() - t0
    return n_patches * 9 * steps / dt, dt / steps * 1e3, threads, f"{steps} steps of {n_patches} patches x 9 rays (of 2048) at the start of the schedule, analytic initial occupancy grid, no grid updates in the timed steps"


def ref_cuda_path_run(dev, warmup, steps, backend="reference"):
    """The reference-shaped step on this GPU (oracle/cuda_path.py): the UNMODIFIED reference nerfacc kernels
    (oracle/_ref) + ATen ops + torch.optim.Adam, with the encoding behind tiny-cuda-nn's calling convention
    (tcnn itself is not in the image).  Same workload, same schedule window [warmup, warmup+steps)."""
    from oracle import cuda_path as cp
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    if backend == "reference" and cp.load_ref_nerfacc() is None:
        return None
    ds = SyntheticDataset(SyntheticScene(), device=dev)
    tr = cp.CudaTrainer(ds, dict(DILIGENT_CONF), backend=backend, seed=0, device=dev)
    ms, spr = cp.time_steps(tr, steps, warmup)
    n = DILIGENT_CONF["batch_size"]
    return {"value": n * 9 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup, "samples_per_ray_last": spr,
            "what": ("unmodified reference nerfacc 0.3.5 kernels (oracle/_ref, sm_100a) + ATen/autograd/torch.optim.Adam step; "
                     "encoding = supernormal_b200.tcnn_api under tcnn's convention (all 14 levels, mask after, fp16 recast per forward) "
                     "because tiny-cuda-nn is absent") if backend == "reference" else
                    "same ATen/autograd step over the drop-in modules supernormal_b200.nerfacc_api + tcnn_api (no fused trainer)"}


_STDOUT_FD = None


def guard_stdout():
    """stdout carries exactly ONE JSON line: everything else any library writes to fd 1 (NCCL's version banner, warnings
    of extensions) is sent to stderr; emit() writes the line to the real stdout."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        os.write(1, data)
    else:
        os.write(_STDOUT_FD, data)


def run_reference(args, rank):
    if rank != 0:
        return
    # each step = a 64-patch sample of the 2048-patch batch (~1 s on 16 cores); K steps, or as many as fit into ~100 s of CPU work
    warm = max(1, min(args.warmup, 3))
    v, ms, threads, sample = cpu_port_run(max(1, args.steps), warm, budget_s=100.0)
    steps = int(sample.split()[0])
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "steps_requested": args.steps,
            "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference has no CPU path (nerfacc/tcnn are CUDA-only); this is the PyTorch-CPU port of its operators (oracle/)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=10 ** 9, help="cap on the e2e arm's steps (default: same K)")
    ap.add_argument("--workload", default="diligent", choices=["diligent", "own_objects"],
                    help="diligent: BASELINE configs[1] (default, the headline); own_objects: configs[2] (36 views 1512x2016, own_objects.conf schedule)")
    ap.add_argument("--flush-l2", action="store_true", help="time every step separately and overwrite a 512 MB buffer between steps (cold L2)")
    ap.add_argument("--ref-cuda-steps", type=int, default=100, help="steps of the reference-shaped CUDA path timed beside ours at N=1 (0 = skip)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    guard_stdout()
    if args.impl == "reference":
        return run_reference(args, rank)
    assert args.warmup >= 3, "timing rules: W >= 3"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    from supernormal_b200 import _lib
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    from supernormal_b200.trainer import FusedTrainer

    from supernormal_b200.synthetic import OWN_OBJECTS_CONF
    if args.workload == "own_objects":
        # view count / resolution of the authors' own captures are not pinned in the reference repo (data not shipped): SURVEY.md §8d
        scene, CONF = SyntheticScene(n_views=36, H=1512, W=2016, exclude_views=()), OWN_OBJECTS_CONF
        workload = "own_objects.conf-shaped training: 36 views 2016x1512 synthetic sphere normals, 2048 patches x 3x3 rays/step, 14-level hash grid T=2^19, 30000-it schedule from random init"
    else:
        scene, CONF, workload = SyntheticScene(), DILIGENT_CONF, WORKLOAD
    ds = SyntheticDataset(scene, device=dev)
    tr = FusedTrainer(ds, dict(CONF), device=dev, seed=0, world_size=world, rank=rank)
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clk = ClockSampler(local)
    if rank == 0:   # rank 0 prints the line; one poller per box
        clk.start()
    for _ in range(W):
        tr.train_step()
    # ---- timed region: device-resident sampling ----------------------------------------------
    kern_ev = []
    _lib.LAUNCH_COUNT = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    spr = []
    barrier()
    clk.mark_start()
    torch.cuda.profiler.start()   # `ncu --profile-from-start off` captures exactly the timed steps
    ev0.record()
    Kr = min(K, args.ref_cuda_steps) if (world == 1 and args.workload == "diligent") else 0
    evr = torch.cuda.Event(enable_timing=True)
    flush_ms = 0.0
    if args.flush_l2:
        junk = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        pairs = []
    for i in range(K):
        if args.flush_l2:
            junk.fill_(i & 0xff)
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        tr.train_step()
        if args.flush_l2:
            b_.record()
            pairs.append((a, b_))
        if i + 1 == Kr:
            evr.record()   # our time over the same schedule window the reference-shaped path is timed on
    ev1.record()
    barrier()
    clk.mark_end()
    torch.cuda.profiler.stop()
    clk.stop()
    ms = ev0.elapsed_time(ev1)
    if args.flush_l2:
        ms = sum(a.elapsed_time(b_) for a, b_ in pairs)   # steps only; the flush writes between them are not counted
    launches = _lib.LAUNCH_COUNT
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = tr.n_patches * 9 * K * world / (ms * 1e-3)
    lt = tr.loss_terms()

    # ---- per-kernel timing of the dominant kernels (CUDA events on the launching stream) ----------
    prof = tr.profile_kernels(steps=min(50, K)) if hasattr(tr, "profile_kernels") else {}

    # ---- e2e: the same schedule window [W, W+Ke) on a fresh trainer, every step's patch batch copied from pinned HOST
    # memory inside the timed region and the loss terms read back every step ------------------------------------
    Ke = min(args.e2e_steps, K)
    pool = [{k: v.cpu() for k, v in tr.sample_batch().items()} for _ in range(16)]     # HOST batches (a CPU loader's output)
    jit_pool = [torch.rand(tr.n_patches) for _ in range(16)]
    te = FusedTrainer(ds, dict(CONF), device=dev, seed=0, world_size=world, rank=rank)
    for _ in range(W):
        te.train_step()
    feeder = te.host_feeder(depth=3, log_capacity=max(Ke, 1))
    h2d, d2h = feeder.h2d_bytes, feeder.d2h_bytes
    packed = [feeder.pack(pool[i], jit_pool[i]) for i in range(16)]   # the loader's output: packed batches in pinned host memory
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    feeder.submit(packed[0])
    for i in range(Ke):
        feeder.step()                                                   # launches step i (+ async D2H of its loss terms)
        if i + 1 < Ke:
            feeder.submit(packed[(i + 1) % 16])                         # H2D of batch i+1 overlaps step i's kernels
    e1.record()
    barrier()
    e2e_losses = feeder.losses()
    assert len(e2e_losses) == Ke and all(np.isfinite(l["loss"]) for l in e2e_losses), "e2e arm produced a non-finite loss"
    ems = e0.elapsed_time(e1)
    t = torch.tensor([ems], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = te.n_patches * 9 * Ke * world / (float(t.item()) * 1e-3)

    if rank == 0:
        peak, which = peaks()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "patches_per_gpu": tr.n_patches, "parallelism": f"dp{world}",
                           "l2_policy": ("flushed: 512 MB overwritten between steps, every step timed separately with CUDA events" if args.flush_l2 else
                                         "not flushed: a training run is a dependent chain of steps -- parameters are rewritten and patches are new random draws every step, "
                                         "so no step repeats an input; the fp16 hash table (<= 24 MB) staying L2-resident across steps is part of the design (--flush-l2 times the cold-L2 case)"),
                           "final_iter": tr.iter_step, "samples_per_ray_last": lt["samples_per_ray"], "loss_last": lt["loss"]},
                "clocks": clk.summary(), "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                        "how": "HostBatchFeeder: packed batch in pinned host memory -> one H2D copy per step on a copy stream (double-buffered), loss terms D2H every step into a pinned ring",
                        "loss_last": e2e_losses[-1]["loss"] if e2e_losses else None},
                "kernels": prof}
        if prof.get("avg_samples"):   # samples/s beside patch-rays/s, because samples per ray change along the schedule
            line["patch_ray_samples_per_s"] = {"value": prof["avg_samples"] * 9 * world / (ms / K * 1e-3), "samples_per_ray": prof["avg_samples"] / tr.n_patches,
                                               "note": "sample count averaged over the profiled steps right after the timed window"}
        if prof.get("dominant"):
            dk = prof["dominant"]
            traffic, traffic_note = None, None
            tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")   # dram__bytes_read+write per launch from one `ncu --set full` capture
            if os.path.exists(tp):
                tj = json.load(open(tp))
                if tj.get("entry_point") == dk["name"]:
                    traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("note")
            line["roofline"] = {"bound": "hbm", "kernel": dk["name"], "achieved": dk["gbs"], "peak": peak, "unit": "GB/s", "frac": dk["gbs"] / peak,
                                "traffic": traffic, "traffic_note": traffic_note, "peak_source": which, "algorithmic_bytes_per_launch": dk["bytes"],
                                "avg_us": dk["us"],
                                "note": "the step's kernels are latency / issue / L2-atomic bound at 2048 patches per GPU (DESIGN.md section 5); the only HBM-streaming "
                                        "kernel is the Adam sweep inside the step-tail kernel (snb_train_tail in kernels.hbm_model)"}
        if Kr > 0:
            ours_ms = ev0.elapsed_time(evr) / Kr
            for backend, key in (("reference", "reference_cuda_path"), ("dropin", "dropin_api_path")):
                try:
                    r = ref_cuda_path_run(dev, W, Kr, backend)
                except Exception as e:   # measurement aid only: never take the bench line down with it
                    r = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
                if r is not None:
                    if "ms_per_step" in r:
                        r["ours_ms_per_step_same_window"] = ours_ms
                    line[key] = r
        if world == 1 and not args.no_cpu:
            v, cms, threads, sample = cpu_port_run(3, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "ms_per_step": cms}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
