"""Per-ray clock64 durations of the marcher (debug build: scripts/build_variant.sh marchdbg -DSNB_MARCH_DEBUG, swapped over lib/libsnb200.so).
usage: python scripts/march_ray_times.py 1000 4800"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from supernormal_b200 import _lib
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(), device=dev)
tr = FusedTrainer(ds, dict(DILIGENT_CONF), device=dev)
fn = _lib.lib().snb_debug_march_stats
fn.restype, fn.argtypes = C.c_int32, [C.c_void_p, C.c_int32]
for at in [int(a) for a in sys.argv[1:]] or [1000, 4800]:
    while tr.iter_step < at:
        tr.train_step()
    torch.cuda.synchronize()
    tr.train_step()
    torch.cuda.synchronize()
    out = np.zeros((tr.n_patches, 4), np.int64)
    assert fn(out.ctypes.data, tr.n_patches) == 0
    cyc, win, bat, ns = out.T
    def q(v): return {"mean": float(v.mean()), "p50": float(np.percentile(v, 50)), "p90": float(np.percentile(v, 90)), "p99": float(np.percentile(v, 99)), "max": float(v.max())}
    e, f = ns == 0, ns > 0
    import heapq
    def makespan(order, slots=148 * 6):
        h = [0] * slots
        heapq.heapify(h)
        end = 0
        for i in order:
            t = heapq.heappop(h) + int(cyc[i])
            end = max(end, t)
            heapq.heappush(h, t)
        return end
    sched = {"in_order": makespan(range(len(cyc))), "longest_first": makespan(np.argsort(-cyc)), "nonempty_first": makespan(np.argsort(e, kind="stable")),
             "sum_over_slots": int(cyc.sum() / (148 * 6)), "max_ray": int(cyc.max())}
    top = np.argsort(-cyc)[:8]
    print(json.dumps({"iter": at, "rays": int(len(cyc)), "empty_share": float(e.mean()), "sum_cycles_M": float(cyc.sum() / 1e6),
                      "cycles_empty": q(cyc[e]), "cycles_nonempty": q(cyc[f]), "windows_empty": q(win[e]), "windows_nonempty": q(win[f]),
                      "batches_nonempty": q(bat[f]), "samples_nonempty": q(ns[f]),
                      "list_scheduling_cycles": sched, "slowest": [{"cycles": int(cyc[i]), "windows": int(win[i]), "batches": int(bat[i]), "samples": int(ns[i])} for i in top],
                      "fit": "cycles ~ a + b windows + c batches: " + str(np.round(np.linalg.lstsq(np.stack([np.ones(len(cyc)), win, bat], 1).astype(float), cyc.astype(float), rcond=None)[0], 1).tolist())}))
