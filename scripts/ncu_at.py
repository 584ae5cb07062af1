"""Train the diligent schedule to iteration ITER unprofiled, then run N steps between cudaProfilerStart/Stop (for
`ncu --profile-from-start off ...`).  usage: python scripts/ncu_at.py ITER [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
at, n = int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 3
ds = SyntheticDataset(SyntheticScene(), device=dev)
tr = FusedTrainer(ds, dict(DILIGENT_CONF), device=dev)
while tr.iter_step < at:
    tr.train_step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(n):
    tr.train_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled iterations", at, "..", tr.iter_step, tr.loss_terms())
