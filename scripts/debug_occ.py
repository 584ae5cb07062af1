import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import torch
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
cuda = torch.device('cuda:0')
ds = SyntheticDataset(SyntheticScene(n_views=4, H=64, W=80, exclude_views=(0,)), device=cuda)
a, b = FusedTrainer(ds, dict(DILIGENT_CONF, batch_size=64), device=cuda), FusedTrainer(ds, dict(DILIGENT_CONF, batch_size=64), device=cuda)
b.fused_host = False
r = torch.arange(128, device=cuda).float().add(0.5).div(64).sub(1)
gx, gy, gz = torch.meshgrid(r, r, r, indexing="ij")
rad = (gx ** 2 + gy ** 2 + gz ** 2).sqrt()
pts = torch.stack([gx, gy, gz], -1).reshape(-1, 3)
a.model.prep()
sdf = a.model.sdf(pts).view(128, 128, 128)
print("sdf at centre-ish", sdf[64, 64, 64].item(), "min", sdf.min().item(), "frac sdf<0.027", (sdf < 0.027).float().mean().item())
for it in (0, 8, 256, 264):
    a.update_occupancy(it); b.update_occupancy(it)
    for name, t in (("fused", a), ("aten", b)):
        bad = (~t.grid.binary) & (rad < 0.5)
        print(it, name, "occupied frac", t.grid.binary.float().mean().item(), "unoccupied inside r<0.5:", int(bad.sum()),
              "their sdf range", (sdf[bad].min().item(), sdf[bad].max().item()) if bad.any() else None,
              "occs there", (t.grid.occs.view(128,128,128)[bad].min().item(), t.grid.occs.view(128,128,128)[bad].max().item()) if bad.any() else None,
              "mean occ", t.grid.occs.mean().item(), "ws", a.occ_ws.view(torch.int64).tolist() if name == "fused" else "")
