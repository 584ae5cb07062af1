"""Is the MAE increase of the 8-GPU run (0.436 deg vs 0.353 deg at N=1, Chamfer equal) a data-parallel defect or the batch size?  One GPU,
the same schedule, 8 x 2048 patches per step: if the single-GPU big batch reproduces the number, it is the optimisation problem that changed
(8x the patches per step at unchanged learning rate and iteration count), not the gradient exchange."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.runner import time_to_mesh
dev = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(), device=dev)
out = {}
for nb in [int(a) for a in sys.argv[1:]] or [2048, 16384]:
    r = time_to_mesh(ds, dict(DILIGENT_CONF, batch_size=nb), 512, device=dev)
    out[nb] = {k: r[k] for k in ("train_s", "chamfer_mm", "fscore", "mae_allview", "mae_testview")}
    print(json.dumps({nb: out[nb]}), flush=True)
