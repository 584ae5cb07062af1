"""What is the host-fed (e2e) step made of?  Times the HostBatchFeeder loop of bench.py on the driver's window with (a) the shipped form, (b) no
per-step loss read-back, (c) no per-step H2D (batches already on the device, same train_step(batch=...) path), (d) the device-sampled step."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer, BATCH_KEYS
dev = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(), device=dev)
W, K = 5, int(sys.argv[1]) if len(sys.argv) > 1 else 200
out = {}


def run(mode):
    tr = FusedTrainer(ds, dict(DILIGENT_CONF), device=dev, seed=0)
    pool = [{k: v.cpu() for k, v in tr.sample_batch().items()} for _ in range(16)]
    jit = [torch.rand(tr.n_patches) for _ in range(16)]
    for _ in range(W):
        tr.train_step()
    fd = tr.host_feeder(depth=3, log_capacity=K)
    packed = [fd.pack(pool[i], jit[i]) for i in range(16)]
    dev_batches = [({k: v.to(dev) for k, v in pool[i].items()}, jit[i].to(dev)) for i in range(16)]
    if mode == "no_d2h":
        fd.log = torch.zeros_like(fd.log)   # un-pinned dummy is never used: patch step() below
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    if mode in ("shipped", "no_d2h"):
        if mode == "no_d2h":
            orig = fd.step
            def step_no_log():
                tr_, slot = fd.tr, fd.n_run % fd.depth
                cur = torch.cuda.current_stream(tr_.device)
                cur.wait_event(fd.ready[slot])
                dv = fd._views(fd.dev[slot])
                tr_.train_step(batch={k: dv[k] for k in BATCH_KEYS}, jitter=dv["jitter"])
                fd.free[slot].record(cur)
                fd.n_run += 1
            fd.step = step_no_log
        fd.submit(packed[0])
        for i in range(K):
            fd.step()
            if i + 1 < K:
                fd.submit(packed[(i + 1) % 16])
    elif mode == "device_batches":
        for i in range(K):
            b, j = dev_batches[i % 16]
            tr.train_step(batch=b, jitter=j)
    else:
        for i in range(K):
            tr.train_step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


for mode in ("shipped", "no_d2h", "device_batches", "device_sampled", "shipped"):
    out.setdefault(mode, []).append(round(run(mode), 5))
print(json.dumps(out))
