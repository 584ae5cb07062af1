"""Full diligent.conf-shaped run on the synthetic sphere: train 5000 it -> 512^3 mesh -> Chamfer / F-score / MAE.
usage (GPU box): python scripts/time_to_mesh.py [--iters 5000] [--res 512] > gpurun_out/time_to_mesh.json"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.runner import time_to_mesh

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=5000)
ap.add_argument("--res", type=int, default=512)
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
ds = SyntheticDataset(SyntheticScene(), device=dev)
conf = dict(DILIGENT_CONF)
if a.iters != 5000:   # compressed schedule: same shape, fewer iterations
    f = a.iters / 5000
    conf.update(end_iter=a.iters, increase_bindwidth_every=max(1, int(350 * f)), warm_up_end=max(1, int(50 * f)))
out = time_to_mesh(ds, conf, a.res, device=dev)
out.pop("vertices", None); out.pop("triangles", None)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
