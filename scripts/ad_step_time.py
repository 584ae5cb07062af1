import sys, time; sys.path.insert(0, '/root/repo')
import torch
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(), device=dev)
tr = FusedTrainer(ds, dict(DILIGENT_CONF, gradient_method="ad"), device=dev)
for _ in range(30): tr.train_step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): tr.train_step()
torch.cuda.synchronize(); print("ad ms/step (it 30-80):", (time.perf_counter() - t0) / 50 * 1e3, tr.loss_terms())
