"""Device time of one fused occupancy update (snb_occgrid_update_fused) at points of the schedule: warm-up sweep (all cells) and sparse sweeps.
usage: python scripts/sweep_time.py 8 1000 4800"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(), device=dev)
tr = FusedTrainer(ds, dict(DILIGENT_CONF), device=dev)
out = {}
for at in [int(a) for a in sys.argv[1:]] or [8, 1000, 4800]:
    while tr.iter_step < at:
        tr.train_step()
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        tr.update_occupancy(at)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    out[at] = {"n_active": tr.model.n_active, "us_min": round(min(ts), 1), "us_median": round(sorted(ts)[2], 1), "occupied": float(tr.grid.binary.float().mean())}
print(json.dumps(out))
