"""Per-kernel device time (CUDA events around each C-ABI call) at chosen points of the diligent schedule.
usage: python scripts/kernel_times.py 1000 3000 4800"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(), device=dev)
tr = FusedTrainer(ds, dict(DILIGENT_CONF), device=dev)
for at in [int(a) for a in sys.argv[1:]] or [1000, 3000]:
    while tr.iter_step < at:
        tr.train_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(64):
        tr.train_step()
    e1.record()
    torch.cuda.synchronize()
    p = tr.profile_kernels(steps=40)
    print(json.dumps({"iter": at, "ms_per_step": e0.elapsed_time(e1) / 64, "n_active": p["n_active"], "avg_samples": p["avg_samples"],
                      "us": p["us_per_step"]}))
