"""Equal-iteration quality check (north star: "final-mesh Chamfer/F-score within 2% of the reference after equal
iterations"): the fused B200 path and the reference-shaped CUDA path (oracle/cuda_path.py backend "reference": the
UNMODIFIED reference nerfacc kernels + ATen/autograd/Adam) train the same diligent.conf-shaped schedule on the same
synthetic sphere; both networks then go through the SAME mesh extraction and metric code.
usage (GPU box): python scripts/quality_vs_reference.py [--iters 5000] [--res 512] > gpurun_out/quality_vs_reference.json"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supernormal_b200 import mesh
from supernormal_b200.runner import time_to_mesh, evaluate_sphere_mesh, eval_mae
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
from oracle import cuda_path as cp

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=5000)
ap.add_argument("--res", type=int, default=512)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
ds = SyntheticDataset(SyntheticScene(), device=dev)
conf = dict(DILIGENT_CONF)
if a.iters != 5000:
    f = a.iters / 5000
    conf.update(end_iter=a.iters, increase_bindwidth_every=max(1, int(350 * f)), warm_up_end=max(1, int(50 * f)))

ours = time_to_mesh(ds, conf, a.res, device=dev)
ours.pop("vertices", None); ours.pop("triangles", None)

ref = cp.CudaTrainer(ds, conf, backend="reference", seed=0, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
skipped = 0
while ref.iter_step < conf["end_iter"]:
    loss, _ = ref.step()
    if loss is None:          # the reference skips an iteration whose rays all miss the grid (models/renderer.py:137-141)
        skipped += 1
        if skipped > 100: break
torch.cuda.synchronize(); t1 = time.perf_counter()
sd = {"sdf_network_fine": {"encoding.params": ref.sdf.encoding.params.detach(), "lin0.bias": ref.sdf.lin0.bias.detach(),
                           "lin0.weight_g": ref.sdf.lin0.weight_g.detach(), "lin0.weight_v": ref.sdf.lin0.weight_v.detach(),
                           "lin1.bias": ref.sdf.lin1.bias.detach(), "lin1.weight_g": ref.sdf.lin1.weight_g.detach(),
                           "lin1.weight_v": ref.sdf.lin1.weight_v.detach()},
      "variance_network_fine": {"variance": ref.dev.variance.detach()}}
tr = FusedTrainer(ds, conf, device=dev)
tr.model.load_reference_state_dict(sd)
tr.model.n_active = min(ref.sdf.bindwidth, tr.model.n_levels)
tr.iter_step = ref.iter_step
tr.grid._binary = ref.renderer.occupancy_grid.binary.clone()
tr.model.prep()
v, t = mesh.extract_geometry(tr.model, ds.object_bbox_min, ds.object_bbox_max, a.res, 0.0)
refm = {"train_s": t1 - t0, "iters": ref.iter_step, "n_vertices": int(v.shape[0]), "n_triangles": int(t.shape[0])}
refm.update(evaluate_sphere_mesh(v, t, float(ds.scene.radius)))
refm.update(eval_mae(tr))
rel = lambda k: abs(ours[k] - refm[k]) / max(abs(refm[k]), 1e-12)
print(json.dumps({"iters": a.iters, "resolution": a.res, "ours_fused": ours, "reference_shaped_cuda_path": refm,
                  "rel_diff": {k: rel(k) for k in ("chamfer_mm", "fscore", "mae_allview")},
                  "speedup_train": refm["train_s"] / ours["train_s"]}))
