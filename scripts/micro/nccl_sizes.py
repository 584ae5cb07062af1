"""NCCL collective timings at the gradient sizes of the diligent schedule (torchrun, one process per GPU).
all-reduce fp32 of the live range vs reduce-scatter fp32 + all-gather fp16 (sharded-optimizer alternative)."""
import os, json, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def timed(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3
out = []
for mb, name in ((1.4e6, "4 levels"), (3.8e6, "8 levels"), (7.7e6, "11 levels"), (11.9e6, "14 levels")):
    n = int(mb) // (8 * world) * (8 * world)
    g = torch.randn(n, device=dev); shard = torch.empty(n // world, device=dev)
    h = torch.empty(n, dtype=torch.float16, device=dev); hs = torch.empty(n // world, dtype=torch.float16, device=dev)
    out.append({"floats": n, "what": name, "allreduce_f32_us": round(timed(lambda: dist.all_reduce(g)), 1),
                "reduce_scatter_f32_us": round(timed(lambda: dist.reduce_scatter_tensor(shard, g)), 1),
                "all_gather_f16_us": round(timed(lambda: dist.all_gather_into_tensor(h, hs)), 1)})
if rank == 0: print(json.dumps(out))
dist.destroy_process_group()
