"""own_objects.conf's val_mesh_res = 1024 (config/own_objects.conf, exp_runner.py:483-506): 1024^3 lattice query + GPU marching cubes on one GPU,
whole volume at once and in 4 sequential x-slabs; vertex / triangle counts must agree, and the 512^3 mesh of the same SDF must describe the same
surface (radius statistics).  usage: python scripts/mesh_1024_check.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from supernormal_b200 import mesh
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(), device=dev)
conf = dict(DILIGENT_CONF, end_iter=1000, increase_bindwidth_every=70, warm_up_end=10)
tr = FusedTrainer(ds, conf, device=dev)
for _ in range(1000):
    tr.train_step()
out = {"n_active": tr.model.n_active}
for res, slabs in ((512, 1), (1024, 1), (1024, 4)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    v, t = mesh.extract_geometry(tr.model, ds.object_bbox_min, ds.object_bbox_max, res, 0.0, slabs=slabs)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    r = np.linalg.norm(v, axis=1)
    out[f"{res}^3 slabs={slabs}"] = {"seconds": dt, "n_vertices": int(v.shape[0]), "n_triangles": int(t.shape[0]), "radius_mean": float(r.mean()), "radius_std": float(r.std()),
                                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
print(json.dumps(out))
