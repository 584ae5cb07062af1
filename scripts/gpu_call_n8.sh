#!/bin/bash
# N = 8 validation of the round-2 state: the driver's scaling command, a longer window, and the replica bit-identity check
set -u
mkdir -p gpurun_out
N=${1:-8}
make -s -C oracle
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu \
    > gpurun_out/bench_n${N}_w5k20.json 2> gpurun_out/bench_n${N}_w5k20.err
echo "bench N=$N rc=$?"; python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_n${N}_w5k20.json"))
    for k in ("value", "ms_per_step", "e2e", "continuation", "time_to_mesh", "schedule_avg", "clocks", "gpu_launches"):
        print(k, json.dumps(j.get(k))[:500])
except Exception as e:
    print("no json", e)
PY
tail -c 1500 gpurun_out/bench_n${N}_w5k20.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 scripts/dp_peer_check.py --every 3 --steps 100 \
    > gpurun_out/dp_peer_check_n${N}.json 2> gpurun_out/dp_peer_check_n${N}.err
echo "dp_peer_check N=$N rc=$?"; grep "^{" gpurun_out/dp_peer_check_n${N}.json | cut -c1-900; tail -c 600 gpurun_out/dp_peer_check_n${N}.err
