import torch, sys
sys.path.insert(0, ".")
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
cuda = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(n_views=6, H=64, W=80, exclude_views=(0,)), device=cuda)
conf = dict(DILIGENT_CONF, batch_size=300, end_iter=200, increase_bindwidth_every=5)
a, b = FusedTrainer(ds, conf, device=cuda), FusedTrainer(ds, conf, device=cuda)
b.fused_host = False
b.legacy_render = True
a.train_step()
b.grid._binary.copy_(a.grid._binary); b.grid.occs.copy_(a.grid.occs); b.update_occupancy = lambda it: None
b.train_step(batch={k: v.clone() for k, v in a.own_batch.items()}, jitter=a.own_jitter.clone())
S = a.buf.totals[0].item(); nS = 9 * S
print("S", S, a.buf.totals.tolist(), b.buf.totals.tolist())
for name in ("d_sdf0", "d_sdf1"):
    x, y = getattr(a.buf, name)[:nS].view(S, 9), getattr(b.buf, name)[:nS].view(S, 9)
    d = (x - y).abs()
    print(name, "norm diff", (x - y).norm().item(), "norm", y.norm().item(), "max", d.max().item(), "n>1e-4", (d > 1e-4).sum().item())
    idx = torch.nonzero(d.max(1).values > 1e-4).flatten()[:12]
    pk = a.buf.packed_info
    for s in idx.tolist():
        p = a.buf.patch_idx[s].item(); base, cnt = pk[p].tolist()
        print("  s", s, "patch", p, "j", s - base, "of", cnt, "end_slot", a.buf.end_slot[s].item(), "a", x[s, :3].tolist(), "b", y[s, :3].tolist())
