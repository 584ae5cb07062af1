"""In-kernel timeline of the data-parallel step tail (train_tail_peer_kernel) on N GPUs: block 0 of every launch stamps %globaltimer at
kernel start | all ranks' "backward complete" flags seen | local reduce + Adam + broadcast fenced | all ranks' "stores done" flags seen.
Per rank and step (each GPU's own clock, so only differences on one rank are used):
  wait_start = t1 - t0   waiting for the slowest rank's backward (rank skew) + one NVLink flag hop
  work       = t2 - t1   peer loads, sharded Adam, fp16 broadcast, system fence
  wait_done  = t3 - t2   waiting for the slowest rank's stores + one flag hop
usage: SNB_PEER_TRACE=4096 python -m torch.distributed.run --nproc-per-node N scripts/dp_peer_trace.py [--warmup 5 --steps 400]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SNB_PEER_TRACE", "4096")
import torch
import torch.distributed as dist
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--steps", type=int, default=400)
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
ds = SyntheticDataset(SyntheticScene(), device=dev)
tr = FusedTrainer(ds, dict(DILIGENT_CONF), device=dev, seed=0, world_size=world, rank=rank)
assert tr.peer_mode and tr.model.peer_trace is not None
for _ in range(a.warmup):
    tr.train_step()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    tr.train_step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
t = tr.model.peer_trace.cpu()
ep = torch.arange(a.warmup + 1, a.warmup + a.steps + 1) % t.shape[0]       # epochs of the timed launches (launch counter starts at 1)
t = t[ep].double()
ws, wk, wd, tot = (t[:, 1] - t[:, 0]) / 1e3, (t[:, 2] - t[:, 1]) / 1e3, (t[:, 3] - t[:, 2]) / 1e3, (t[:, 3] - t[:, 0]) / 1e3
gap = (t[1:, 0] - t[:-1, 3]) / 1e3                                         # end of one tail -> start of the next = the rest of the step on this rank
stat = lambda x: [round(float(x.mean()), 2), round(float(x.median()), 2), round(float(x.quantile(0.95)), 2)]
mine = {"rank": rank, "ms_per_step": ms, "wait_start_us": stat(ws), "work_us": stat(wk), "wait_done_us": stat(wd), "tail_total_us": stat(tot), "rest_of_step_us": stat(gap)}
allr = [None] * world
dist.all_gather_object(allr, mine)
if rank == 0:
    print(json.dumps({"world": world, "steps": a.steps, "iterations": [a.warmup, a.warmup + a.steps], "stat": "[mean, median, p95] in us", "ranks": allr}))
dist.destroy_process_group()
