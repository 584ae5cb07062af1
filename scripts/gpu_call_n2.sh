#!/bin/bash
# N = 2 visit: the 2-rank peer-tail test of the suite, the driver's scaling command at N = 2, replica bit-identity check
set -u
mkdir -p gpurun_out
make -s -C oracle
timeout 300 python -m pytest tests/test_gpu_fused.py -q -m gpu -k "peer_tail" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu \
    > gpurun_out/bench_n2_w5k20.json 2> gpurun_out/bench_n2_w5k20.err
echo "bench N=2 rc=$?"; python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_n2_w5k20.json"))
    for k in ("value", "ms_per_step", "e2e", "continuation", "time_to_mesh", "schedule_avg", "clocks", "gpu_launches"):
        print(k, json.dumps(j.get(k))[:400])
except Exception as e:
    print("no json", e)
PY
tail -c 800 gpurun_out/bench_n2_w5k20.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 scripts/dp_peer_check.py --every 3 --steps 100 \
    > gpurun_out/dp_peer_check_n2.json 2> gpurun_out/dp_peer_check_n2.err
echo "dp_peer_check N=2 rc=$?"; grep "^{" gpurun_out/dp_peer_check_n2.json | cut -c1-700; tail -c 400 gpurun_out/dp_peer_check_n2.err
