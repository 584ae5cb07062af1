"""Host (Python + launch) time per training step against the device time: is the loop GPU-bound?  Device-sampled and host-fed paths.
usage: python scripts/host_time_per_step.py [steps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(), device=dev)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 300
out = {}
for mode in ("device_sampled", "host_fed"):
    tr = FusedTrainer(ds, dict(DILIGENT_CONF), device=dev, seed=0)
    pool = [{k: v.cpu() for k, v in tr.sample_batch().items()} for _ in range(16)]
    jit = [torch.rand(tr.n_patches) for _ in range(16)]
    for _ in range(5):
        tr.train_step()
    fd = tr.host_feeder(depth=3, log_capacity=K)
    packed = [fd.pack(pool[i], jit[i]) for i in range(16)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    if mode == "device_sampled":
        for _ in range(K):
            tr.train_step()
    else:
        fd.submit(packed[0])
        for i in range(K):
            fd.step()
            if i + 1 < K:
                fd.submit(packed[(i + 1) % 16])
    e1.record()
    t_host = time.perf_counter() - t0          # the loop has only ENQUEUED the work here
    torch.cuda.synchronize()
    out[mode] = {"host_us_per_step": round(t_host / K * 1e6, 1), "device_us_per_step": round(e0.elapsed_time(e1) / K * 1e3, 1)}
print(json.dumps(out))
