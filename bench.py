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


L2_RED_PEAK_GOPS = 195.2   # G reduction sectors/s, measured on this pool's B200 (profiles/r02_red_rate_microbench.txt)


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


CPU_CONFIG = ("BASELINE configs[0]: synthetic sphere normal maps, 20 views 612x512, 3x3 patches, 512 patches/batch, 16-level hash grid "
              "(T=2^19, base 32), one training step fwd+bwd+Adam via the PyTorch-CPU expression of the operators (oracle/)")


def cpu_port_run(steps=8, warmup=2, n_patches=512, n_levels=16, threads=None, budget_s=None):
    """The oracle (PyTorch-CPU port of the reference operators) on BASELINE.json configs[0] as written: 512 patches x 9
    rays, 16 levels, real occupancy grid with its every-8-iterations update INSIDE the timed steps (warm-ups are iterations 0-1, so with
    >= 7 timed steps iteration 8 carries one update: exactly the 1-in-8 amortisation).  -> (patch-rays/s from the mean step, mean ms,
    median ms, threads, sample description)."""
    from oracle import torch_ops as T
    from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    ds = SyntheticDataset(SyntheticScene(), device="cpu")
    conf = dict(DILIGENT_CONF, batch_size=n_patches, encoding=dict(DILIGENT_CONF["encoding"], n_levels=n_levels))
    tr = T.Trainer(ds, conf, seed=0)
    for _ in range(warmup):
        tr.step()
    times = []
    t_all = time.perf_counter()
    while len(times) < steps:
        t0 = time.perf_counter()
        tr.step()
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_all > budget_s:   # bounded sample: stop after the step that crosses the budget
            break
    mean_s, med_s = float(np.mean(times)), float(np.median(times))
    n_upd = sum(1 for it in range(warmup, warmup + len(times)) if it % conf["ray_marching"]["occ_update_freq"] == 0)
    sample = (f"{len(times)} steps (iterations {warmup}..{warmup + len(times) - 1} after {warmup} warm-ups) of {n_patches} patches x 9 rays, {n_levels} levels, "
              f"{n_upd} occupancy-grid update(s) inside the timed steps; value from the mean step, median {med_s * 1e3:.0f} ms")
    return n_patches * 9 / mean_s, mean_s * 1e3, med_s * 1e3, threads, sample


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
    """`--impl reference`: the reference has no CPU path of its own (nerfacc / tiny-cuda-nn are CUDA-only and tiny-cuda-nn is absent from the
    image), so this arm times the oracle port on the host cores: K steps of BASELINE configs[0] (512 patches, 16 levels; ~1.5-3 s each),
    or as many as fit into ~100 s."""
    if rank != 0:
        return
    warm = max(2, min(args.warmup, 2))
    v, ms, med, threads, sample = cpu_port_run(max(1, args.steps), warm, budget_s=100.0)
    steps = int(sample.split()[0])
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "steps_requested": args.steps,
            "warmup": warm,
            "ms_per_step": ms, "ms_per_step_median": med, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CPU_CONFIG, "note": "reference has no CPU path (nerfacc/tcnn are CUDA-only); this is the PyTorch-CPU port of its operators (oracle/); "
                       "patch-rays/s is per patch ray, so the 512-patch CPU sample and the 2048-patch GPU batch are comparable per unit of work"},
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
    ap.add_argument("--no-ttm", action="store_true", help="skip the time-to-mesh leg (full schedule + 512^3 extraction on a fresh trainer)")
    ap.add_argument("--cont-steps", type=int, default=400, help="secondary timed window: this many further iterations right after the K timed ones (0 = skip)")
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
    # ---- secondary window: the schedule simply continues for a few hundred more iterations (a K of 20 is a 5 ms region) ----
    cont = None
    if args.cont_steps > 0 and not args.flush_l2:
        clk2 = ClockSampler(local)
        if rank == 0:
            clk2.start()
            time.sleep(0.05)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        clk2.mark_start()
        c0.record()
        for _ in range(args.cont_steps):
            tr.train_step()
        c1.record()
        barrier()
        clk2.mark_end()
        clk2.stop()
        tc = torch.tensor([c0.elapsed_time(c1)], device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        cms = float(tc.item())
        cont = {"iterations": [W + K, W + K + args.cont_steps], "steps": args.cont_steps, "ms_per_step": cms / args.cont_steps,
                "value": tr.n_patches * 9 * args.cont_steps * world / (cms * 1e-3), "unit": UNIT, "clocks": clk2.summary() if rank == 0 else None}

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

    # ---- time-to-mesh (BASELINE.json metric, second half): the FULL schedule from random init on a fresh trainer + 512^3 extraction
    # (slab-sharded over the ranks), and the schedule-average throughput that goes with it -------------------------------------
    ttm = None
    if not args.no_ttm:
        from supernormal_b200.runner import time_to_mesh
        del te, feeder
        torch.cuda.empty_cache()
        barrier()
        r = time_to_mesh(ds, dict(CONF), 512, device=dev, seed=0, evaluate=(args.workload == "diligent"))
        r.pop("vertices", None)
        r.pop("triangles", None)
        tt = torch.tensor([r["train_s"], r["time_to_mesh_s"]], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        r["train_s"], r["time_to_mesh_s"] = float(tt[0]), float(tt[1])
        ttm = r

    if rank == 0:
        peak, which = peaks()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (TF32 dz in backward)",
                "data": "synthetic",
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
        if cont:
            line["continuation"] = cont
        if ttm:
            n_it = ttm["iters"]
            line["time_to_mesh"] = {"value": ttm["time_to_mesh_s"], "unit": "s", "train_s": ttm["train_s"], "mesh_s": ttm["mesh_s"], "iters": n_it,
                                    "mesh_resolution": ttm["resolution"], "n_vertices": ttm.get("n_vertices"), "n_triangles": ttm.get("n_triangles"),
                                    "what": "wall clock from the first iteration of a fresh trainer to the mesh (vertices / faces) in host memory: end_iter training "
                                            "iterations + extract_geometry(512), slab-sharded over the ranks",
                                    **{k: ttm[k] for k in ("chamfer_mm", "fscore", "radius_mean", "mae_allview", "mae_testview") if k in ttm}}
            line["schedule_avg"] = {"value": tr.n_patches * 9 * n_it * world / ttm["train_s"], "unit": UNIT, "ms_per_step": ttm["train_s"] / n_it * 1e3,
                                    "what": "whole-schedule average (all levels come live along the way); host wall clock around the full training loop"}
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
            # second yardstick for the same kernel: its table scatter is bound by the L2 reduction path, not by HBM.  Peak = measured
            # red.global.add.v2.f32 rate into an L2-resident 34 MB table (scripts/micro/red_rate.cu, profiles/r02_red_rate_microbench.txt:
            # 97-100 reduction sectors per clock chip-wide, whatever the operand width).  Achieved = ALGORITHMIC corner updates (8 per point
            # and live level) per second; the kernel merges same-cell neighbours before it issues, so the issued sector count is lower.
            hm = prof.get("hbm_model", {}).get(dk["name"], {})
            if hm.get("l2_useful_bytes"):
                upd = hm["l2_useful_bytes"] / 8.0           # 8 B (one float2) per corner update
                gops = upd / (dk["us"] * 1e-6) / 1e9
                line["roofline_l2_reduction"] = {"bound": "l2-reduction", "kernel": dk["name"], "achieved": gops, "peak": L2_RED_PEAK_GOPS, "unit": "G corner updates/s vs G red sectors/s",
                                                 "frac": gops / L2_RED_PEAK_GOPS, "corner_updates_per_launch": upd, "avg_us": dk["us"],
                                                 "peak_source": "measured: profiles/r02_red_rate_microbench.txt (red.v2.f32, 33.6 MB table, 32 warps/SM)"}
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
            v, cms, cmed, threads, sample = cpu_port_run(8, 2)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "ms_per_step": cms,
                                    "ms_per_step_median": cmed, "config": CPU_CONFIG}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
