"""Data-parallel check on N GPUs of one box: the peer-memory step tail (snb_train_tail_peer: reduction over NVLink peer loads,
sharded Adam, parameter broadcast -- one kernel) against the NCCL path (snb_unfold_grads -> all_reduce -> snb_train_tail).
  * replicas stay bit-identical on every rank (parameters, fp16 table, folded net) in both modes;
  * both modes train alike from the same initial state (same batches; gradients differ by fp32 summation order only);
  * ms per step of both modes on the same schedule window.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/dp_peer_check.py [--steps 200]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--check-steps", type=int, default=24)
ap.add_argument("--every", type=int, default=350, help="increase_bindwidth_every (smaller: more live levels in the check)")
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
ds = SyntheticDataset(SyntheticScene(), device=dev)
conf = dict(DILIGENT_CONF, increase_bindwidth_every=a.every)


def make(mode):
    os.environ["SNB_DP"] = mode
    return FusedTrainer(ds, dict(conf), device=dev, seed=0, world_size=world, rank=rank)


def replicas_identical(tr):
    """every rank holds the same bits"""
    m = tr.model
    flat = torch.cat([m.flat[:m.n_small], m.gather_table()])   # gather_table: collective only with SNB_PEER_F16ONLY=1
    sig = torch.stack([flat.view(torch.int32).long().sum(), m.table_f16.view(torch.int16).long().sum(), m.net.view(torch.int32).long().sum(),
                       flat.view(torch.int32).long().mul(torch.arange(flat.numel(), device=dev) % 8191).sum()])
    all_sig = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(all_sig, sig)
    return all(torch.equal(all_sig[0], s) for s in all_sig)


out = {"world": world}
A, B = make("peer"), make("nccl")
out["peer_mode"] = [A.peer_mode, B.peer_mode]
out["peer_f16_only"] = bool(getattr(A.model, "peer_f16_only", False))
ident, close, losses = [], [], []
for it in range(a.check_steps):
    A.train_step()
    B.train_step()
    if it in (0, 1, a.check_steps - 1):
        ident.append((replicas_identical(A), replicas_identical(B)))
        close.append(((A.model.gather_table() - B.model.gather_table()).abs() > 1e-6).float().mean().item())
    la, lb = A.loss_terms(), B.loss_terms()
    losses.append((la["loss"], lb["loss"], la["n_samples"], lb["n_samples"]))
out["replicas_identical_peer_nccl"] = ident
out["frac_params_apart_gt_1e-6"] = close
out["loss_first_last"] = [losses[0], losses[-1]]
out["loss_max_rel_diff"] = max(abs(x - y) / max(abs(y), 1e-9) for x, y, _, _ in losses)
out["grad_zero_after_step"] = bool((A.model.grad == 0).all().item())
out["n_active"] = A.model.n_active


def time_steps(tr, n):
    for _ in range(20):
        tr.train_step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        tr.train_step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


out["ms_per_step_peer"] = time_steps(A, a.steps)
out["ms_per_step_nccl"] = time_steps(B, a.steps)
A.check_peer_error()
out["replicas_identical_end"] = (replicas_identical(A), replicas_identical(B))
out["loss_end"] = (A.loss_terms()["loss"], B.loss_terms()["loss"])
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
