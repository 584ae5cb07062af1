"""Per-phase clock64 cycles of sdf_bwd_patch_umma_kernel<8> (thread 0 of every CTA, summed over tiles) on a -DSNB_BWD_DEBUG build.
usage: python scripts/bwd_phase_times.py 15 1000"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from supernormal_b200 import _lib
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(), device=dev)
tr = FusedTrainer(ds, dict(DILIGENT_CONF), device=dev)
fn = _lib.lib().snb_debug_bwd_phases
fn.restype, fn.argtypes = C.c_int32, [C.c_void_p, C.c_int32]
names = ["stage", "barrier 1", "MMA Z + wait", "elementwise + dW1 reduce", "barrier 2", "MMA U issue + dW0 mma.sync", "wait U", "scatter", "barrier 3", "next fetch"]
for at in [int(a) for a in sys.argv[1:]] or [15, 1000]:
    while tr.iter_step < at:
        tr.train_step()
    torch.cuda.synchronize()
    buf = np.zeros(16, np.uint64)
    fn(buf.ctypes.data, 1)
    for _ in range(8):
        tr.train_step()
    torch.cuda.synchronize()
    fn(buf.ctypes.data, 1)
    tiles = int(buf[15])
    tot = float(buf[:10].sum())
    print(json.dumps({"iter": at, "n_active": tr.model.n_active, "tiles_per_step": tiles / 8, "cycles_per_tile": round(tot / max(tiles, 1)),
                      "phases_cycles_per_tile": {n: round(float(buf[i]) / max(tiles, 1)) for i, n in enumerate(names)},
                      "share": {n: round(float(buf[i]) / tot, 3) for i, n in enumerate(names)}}))
