#!/usr/bin/env python
"""BASELINE.json configs[3]: hash-grid encode + SDF MLP + analytic-normal microbenchmark.
2^24 points U(-1,1)^3 (seed 0), 16 levels, F=2, base 32, per-level scale 2^0.4, T in {2^19, 2^22}  (SURVEY.md §8d).

Kernels timed (CUDA events on the launching stream, median of `--reps` after warm-up; the 201 MB point set is larger
than L2, so every repetition re-reads it from HBM):
  encode        snb_hashgrid_fwd          (drop-in tcnn.Encoding forward: f16[N,32] out)
  sdf           snb_sdf_eval              (encode + MLP fused, f32[N] out)
  sdf+normal    snb_sdf_eval_grad         (encode + MLP + analytic d sdf/dx in one pass, f32[N] + f32[N,3] out)
Algorithmic bytes per point: HBM 12 B in + outputs; gathers 16 levels x 8 corners x 4 B = 512 B/pt, which are
L2 traffic for T=2^19 (28 MB fp16 table) and HBM traffic for T=2^22 (183 MB table > 126 MB L2).
Prints one JSON object; `python scripts/microbench_sdf.py --log2-n 24`.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from supernormal_b200 import _lib
from supernormal_b200._lib import call, ptr
from supernormal_b200.trainer import SDFModel


def timed(fn, reps, warmup=2):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-n", type=int, default=24)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--tables", type=int, nargs="*", default=[19, 22])
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    n = 1 << args.log2_n
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(n, 3, device=dev, generator=g) * 2 - 1
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    out = {"n_points": n, "hbm_peak_gbs": peak, "cases": []}
    for log2_t in args.tables:
        enc = dict(otype="HashGrid", n_levels=16, n_features_per_level=2, log2_hashmap_size=log2_t, base_resolution=32,
                   per_level_scale=2 ** 0.4)
        m = SDFModel(enc, device=dev)
        with torch.no_grad():   # non-degenerate features / weights (the init leaves the feature columns of W0 at zero)
            m.flat[: m.n_small].copy_(m.flat[: m.n_small] + 0.05 * torch.randn(m.n_small, device=dev, generator=g))
            m.table.uniform_(-0.05, 0.05, generator=g)
        m.refresh_table_f16()
        m.n_active = 16
        m.prep()
        net = m.net_struct()
        feats = torch.empty(n, 32, dtype=torch.float16, device=dev)
        sdf = torch.empty(n, device=dev)
        grad = torch.empty(n, 3, device=dev)
        meta = C.byref(m.meta)
        runs = {
            "encode": (lambda: call("snb_hashgrid_fwd", n, ptr(x), ptr(m.table_f16), meta, 16, ptr(feats), 0), 12 + 64),
            "sdf": (lambda: call("snb_sdf_eval", n, ptr(x), C.byref(net), 0, ptr(sdf)), 12 + 4),
            "sdf+normal": (lambda: call("snb_sdf_eval_grad", n, ptr(x), C.byref(net), ptr(sdf), ptr(grad)), 12 + 16),
        }
        table_mb = m.n_table * 2 / 1e6
        in_l2 = table_mb < 100
        for name, (fn, hbm_b) in runs.items():
            ms = timed(fn, args.reps)
            alg = hbm_b + (0 if in_l2 else 512)
            out["cases"].append({"log2_T": log2_t, "table_mb_f16": round(table_mb, 1), "kernel": name, "ms": round(ms, 3),
                                 "gpoints_per_s": round(n / ms / 1e6, 3), "hbm_bytes_per_point_algorithmic": alg,
                                 "hbm_gbs_algorithmic": round(n * alg / ms / 1e6, 1), "hbm_frac_of_peak": round(n * alg / ms / 1e6 / peak, 4),
                                 "gather_gbs_useful(512B/pt)": round(n * 512 / ms / 1e6, 1)})
        del m, feats, sdf, grad
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
