"""gradient_method = 'ad' step time on the diligent workload: the fused step (default) and the autograd route (SNB_AD_FUSED=0).
usage: python scripts/ad_step_time.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(), device=dev)
out = {}
for fused, windows in ((1, (30, 1000, 4800)), (0, (30,))):
    os.environ["SNB_AD_FUSED"] = str(fused)
    tr = FusedTrainer(ds, dict(DILIGENT_CONF, gradient_method="ad"), device=dev)
    for at in windows:
        while tr.iter_step < at:
            tr.train_step()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(50):
            tr.train_step()
        torch.cuda.synchronize()
        out[f"fused={fused} it {at}-{at + 50}"] = {"ms_per_step": (time.perf_counter() - t0) / 50 * 1e3, "n_active": tr.model.n_active, **{k: tr.loss_terms()[k] for k in ("loss", "samples_per_ray")}}
    if fused:
        p = tr.profile_kernels(steps=20)
        out["fused kernels us/step at it ~4850"] = p["us_per_step"]
print(json.dumps(out))
