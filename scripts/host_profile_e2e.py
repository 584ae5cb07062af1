"""cProfile of the host side of the host-fed loop (HostBatchFeeder.step / submit) and of the device-sampled loop: where do the microseconds per step go?"""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from supernormal_b200.synthetic import SyntheticDataset, SyntheticScene, DILIGENT_CONF
from supernormal_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
ds = SyntheticDataset(SyntheticScene(), device=dev)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 400
for mode in ("host_fed", "device_sampled"):
    tr = FusedTrainer(ds, dict(DILIGENT_CONF), device=dev, seed=0)
    pool = [{k: v.cpu() for k, v in tr.sample_batch().items()} for _ in range(16)]
    jit = [torch.rand(tr.n_patches) for _ in range(16)]
    for _ in range(5):
        tr.train_step()
    fd = tr.host_feeder(depth=3, log_capacity=K)
    packed = [fd.pack(pool[i], jit[i]) for i in range(16)]
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    if mode == "host_fed":
        fd.submit(packed[0])
        for i in range(K):
            fd.step()
            if i + 1 < K:
                fd.submit(packed[(i + 1) % 16])
    else:
        for _ in range(K):
            tr.train_step()
    pr.disable()
    t = time.perf_counter() - t0
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
    print("=====", mode, "host us/step (under cProfile)", round(t / K * 1e6, 1))
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:32]))
