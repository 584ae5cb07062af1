"""Histogram of the per-patch sample counts at points of the diligent schedule (what a lane mapping of the render stage sees).
usage: python scripts/sample_count_hist.py 15 300 1000 3000 4800"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(), device=dev)
tr = FusedTrainer(ds, dict(DILIGENT_CONF), device=dev)
for at in [int(a) for a in sys.argv[1:]] or [15, 1000, 4800]:
    while tr.iter_step < at:
        tr.train_step()
    acc = torch.zeros(9, dtype=torch.int64)
    edges = torch.tensor([0, 1, 9, 17, 25, 33, 49, 65, 129, 100000])
    tot = 0
    for _ in range(8):
        tr.train_step()
        c = tr.buf.counts.cpu().long()
        tot += int(c.sum())
        acc += torch.histogram(c.float(), bins=edges.float()).hist.long()
    print(json.dumps({"iter": at, "mean": tot / 8 / c.numel(),
                      "share_of_patches": {f"{int(edges[i])}..{int(edges[i + 1]) - 1}": round(float(acc[i]) / float(acc.sum()), 3) for i in range(9)}}))
